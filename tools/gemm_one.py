"""tools/gemm_one.py — one GEMM shape through npm_gemm (for ncu captures / NPM_GEMM_DEBUG_TIMES).
usage: python tools/gemm_one.py <majors> M N K [reps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'np-modeling_b200'))
from npm_b200._lib import C, GemmDesc  # noqa: E402

mj = sys.argv[1]
M, N, K = (int(v) for v in sys.argv[2:5])
reps = int(sys.argv[5]) if len(sys.argv) > 5 and sys.argv[5].isdigit() else 3
a = torch.randn(M, K, device='cuda') if mj[0] == 'k' else torch.randn(K, M, device='cuda')
b = torch.randn(N, K, device='cuda') if mj[1] == 'k' else torch.randn(K, N, device='cuda')
c = torch.empty(M, N, device='cuda')
d = GemmDesc()
bias = torch.zeros(N, device='cuda') if '--bias' in sys.argv else None
d.a, d.b, d.c, d.bias = a.data_ptr(), b.data_ptr(), c.data_ptr(), (bias.data_ptr() if bias is not None else None)
d.m, d.n, d.k = M, N, K
d.a_rs, d.a_cs = (K, 1) if mj[0] == 'k' else (1, M)
d.b_rs, d.b_cs = (1, K) if mj[1] == 'k' else (N, 1)
d.ldc = N
d.nb1, d.nb2 = 1, 1
d.alpha, d.flags, d.precision = 1.0, 0, 0
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
ts = []
for _ in range(reps):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); C.npm_gemm(d, st); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print(f'{mj} M={M} N={N} K={K}: ' + ' '.join(f'{t * 1e3:.1f}us' for t in ts) + f'  best {2.0 * M * N * K / min(ts) / 1e9:.1f} TF', flush=True)
