"""tools/step_trace.py — in-situ kernel timeline of training steps (CUPTI through torch.profiler; no ncu replay, warm
caches, real overlap).  Prints per-kernel totals for one step, the GPU idle time between kernels and the step span.
usage: python tools/step_trace.py [layers] [steps]"""
import collections
import os
import re
import sys

import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'np-modeling_b200'))
import npm_b200  # noqa: E402
import loss as loss_mod  # noqa: E402
import optimizer as opt_mod  # noqa: E402
from layers import adapters  # noqa: E402
from layers.normalizations import set_dropout_seed  # noqa: E402
from npm_b200 import device  # noqa: E402
from train import Trainer, iter_parameters  # noqa: E402

L = int(sys.argv[1]) if len(sys.argv) > 1 else 4
STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 2
B, S, D, H, F = 8, 1024, 1024, 16, 4096
npm_b200.set_precision(os.environ.get('NPM_TRACE_PREC', 'bf16x3'))
np.random.seed(0)
set_dropout_seed(1234)
stack = adapters.DecoderStack(L, H, F, True, 0.1)
trainer = Trainer([stack], loss_mod.MSELoss(), verbose=False, shard_inputs=False)
g = torch.Generator(device='cuda').manual_seed(100)
q, kv, t = (device.DeviceArray(torch.randn(B, S, D, generator=g, device='cuda')) for _ in range(3))
stack(q, kv)
gw = torch.Generator(device='cuda').manual_seed(7)
for owner, name in iter_parameters(stack):
    p = owner._p(name).t
    if name.startswith('_w'):
        fan_in = p.shape[-1] if name in ('_wq', '_wk', '_wv') else (p.shape[1] * p.shape[2] if name == '_wo' else p.shape[0])
        p.copy_(torch.randn(p.shape, generator=gw, device='cuda') / fan_in ** 0.5)
    elif name == '_gamma':
        p.fill_(1.0)
    elif name == '_beta':
        p.zero_()
    else:
        p.copy_(torch.randn(p.shape, generator=gw, device='cuda') * 0.02)
    owner._p(name).touched()
adam = opt_mod.AdamOptimizer(learning_rate=1e-4)
for _ in range(3):
    trainer.train((q, kv), t, 1, adam)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(STEPS):
        trainer.train((q, kv), t, 1, adam)
    torch.cuda.synchronize()
ev = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA),
            key=lambda e: e.time_range.start)
ev = [e for e in ev if 'memcpy' not in e.name.lower() and 'memset' not in e.name.lower() or True]
# last step = events after the second-to-last optimizer kernel
opt_idx = [i for i, e in enumerate(ev) if 'opt_multi' in e.name]
a = opt_idx[-2] + 1 if len(opt_idx) >= 2 else 0
step = ev[a:opt_idx[-1] + 1]
span = step[-1].time_range.end - step[0].time_range.start
busy = sum(e.time_range.end - e.time_range.start for e in step)
gaps = sum(max(0, step[i + 1].time_range.start - step[i].time_range.end) for i in range(len(step) - 1))
agg = collections.defaultdict(lambda: [0, 0.0])
for e in step:
    name = e.name.replace('void ', '').replace('npm::<unnamed>::', '').replace('npm::(anonymous namespace)::', '')
    name = re.sub(r'\(.*', '', name)
    agg[name][0] += 1
    agg[name][1] += e.time_range.end - e.time_range.start
print(f'# {L} layers: one step = {len(step)} GPU activities, span {span / 1e3:.3f} ms, kernel time {busy / 1e3:.3f} ms, '
      f'idle between kernels {gaps / 1e3:.3f} ms ({100 * gaps / span:.1f} %)')
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'{v:10.1f} us {100 * v / span:5.1f}%  n={n:4d}  avg {v / n:7.1f} us  {k[:90]}')
