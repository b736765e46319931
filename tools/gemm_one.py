"""tools/gemm_one.py — the bench.py roofline kernel alone (linear_fwd M=8192 K=1024 N=4096, tf32), for ncu captures."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'np-modeling_b200'))
import npm_b200  # noqa: E402
from npm_b200._lib import C  # noqa: E402
npm_b200.set_precision('tf32')
M, K, N = 8192, 1024, 4096
x = torch.randn(M, K, device='cuda'); w = torch.randn(K, N, device='cuda') / 32; b = torch.zeros(N, device='cuda'); y = torch.empty(M, N, device='cuda')
st = torch.cuda.current_stream().cuda_stream
for _ in range(4):
    C.npm_linear_fwd(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), M, K, N, 0, 0, st)
torch.cuda.synchronize()
ref = (x[:64].double() @ w.double())
print('max rel err', float(((y[:64].double() - ref).abs().max() / ref.abs().max())))
