"""tools/host_overhead.py — how long the HOST takes to enqueue one cfg5 training step (no sync inside)
vs how long the GPU takes to execute it.  If the two are close the step is launch-bound."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'np-modeling_b200')); sys.path.insert(0, ROOT)
import numpy as np, torch
import npm_b200, loss, optimizer
from layers import adapters
from npm_b200 import device
from train import Trainer
L = int(sys.argv[1]) if len(sys.argv) > 1 else 24
npm_b200.set_precision('tf32')
B, S, D = 8, 1024, 1024
g = torch.Generator(device='cuda').manual_seed(0)
q, kv, t = (device.DeviceArray(torch.randn(B, S, D, generator=g, device='cuda')) for _ in range(3))
stack = adapters.DecoderStack(L, 16, 4096, True, 0.1)
tr = Trainer([stack], loss.MSELoss(), verbose=False, shard_inputs=False)
adam = optimizer.AdamOptimizer(learning_rate=1e-4)
for _ in range(3):
    tr.train((q, kv), t, 1, adam)
torch.cuda.synchronize()
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    tr.train((q, kv), t, 1, adam)
    e1.record(); t1 = time.perf_counter()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f'layers={L} host enqueue {1e3 * (t1 - t0):.1f} ms | gpu {e0.elapsed_time(e1):.1f} ms | wall incl. sync {1e3 * (t2 - t0):.1f} ms | launches {npm_b200.launch_count()}')
    npm_b200.reset_launch_count()
