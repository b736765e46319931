"""tools/ffn_planes_probe.py — the FFN GEMMs / ReLU backward of cfg5 with the hidden activation and its gradient as fp32
versus as split-bf16 planes (device.PlanesArray route of layers/mlp.py), one kernel at a time (CUDA events, L2 flushed)."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'np-modeling_b200'))
from npm_b200._lib import C, GemmDesc, load  # noqa: E402

M, D, F = 8192, 1024, 4096
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
st = torch.cuda.current_stream().cuda_stream


def t(fn, iters=8):
    fn(); fn(); ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort(); return ts[len(ts) // 2] * 1e3


def gemm(**kw):
    d = GemmDesc(nb1=1, nb2=1, alpha=1.0, precision=3, **kw)
    rc = load().npm_gemm(ctypes.byref(d), st)
    assert rc == 0, rc


def planes_of(x):
    p = torch.empty(x.numel() * 4, dtype=torch.uint8, device='cuda')
    C.npm_weight_split(x.data_ptr(), p.data_ptr(), x.shape[0], x.shape[1], st)
    return p


g = torch.Generator(device='cuda').manual_seed(0)
x = torch.randn(M, D, device='cuda', generator=g)
w1 = torch.randn(D, F, device='cuda', generator=g) * 0.03
w2 = torch.randn(F, D, device='cuda', generator=g) * 0.03
b1 = torch.randn(F, device='cuda', generator=g)
h = torch.empty(M, F, device='cuda'); hp = torch.empty(M * F * 4, dtype=torch.uint8, device='cuda')
y = torch.empty(M, D, device='cuda')
dyo = torch.randn(M, D, device='cuda', generator=g)
dh = torch.randn(M, F, device='cuda', generator=g)
dz = torch.empty(M, F, device='cuda'); dzp = torch.empty(M * F * 4, dtype=torch.uint8, device='cuda')
dw = torch.empty(F, D, device='cuda'); dw1 = torch.empty(D, F, device='cuda'); dx = torch.empty(M, D, device='cuda')
db = torch.empty(F, device='cuda')
ws = torch.empty(C.npm_colsum_workspace(M, F), dtype=torch.uint8, device='cuda')
w1p, w2p = planes_of(w1), planes_of(w2)
P = lambda a: a.data_ptr()
up = dict(a=P(x), b=P(w1), bias=P(b1), m=M, n=F, k=D, a_rs=D, a_cs=1, b_rs=F, b_cs=1, ldc=F, flags=1, b_split=P(w1p), b_split_plane=w1.numel())
print(f'FFN up   fp32 out   {t(lambda: gemm(c=P(h), **up)):7.1f} us')
print(f'FFN up   planes out {t(lambda: gemm(c=None, c_split=P(hp), c_split_plane=M * F, **up)):7.1f} us')
dn = dict(b=P(w2), c=P(y), m=M, n=D, k=F, a_rs=F, a_cs=1, b_rs=D, b_cs=1, ldc=D, flags=0, b_split=P(w2p), b_split_plane=w2.numel())
print(f'FFN down A fp32     {t(lambda: gemm(a=P(h), **dn)):7.1f} us')
print(f'FFN down A planes   {t(lambda: gemm(a=None, a_split=P(hp), a_split_plane=M * F, **dn)):7.1f} us')
dwd = dict(c=P(dw), m=F, n=D, k=M, a_rs=1, a_cs=F, b=P(dyo), b_rs=D, b_cs=1, ldc=D, flags=0)
print(f'dW2 (h^T dy) A fp32   {t(lambda: gemm(a=P(h), **dwd)):7.1f} us')
print(f'dW2 (h^T dy) A planes {t(lambda: gemm(a=None, a_split=P(hp), a_split_plane=M * F, **dwd)):7.1f} us')
print(f'relu_bwd+colsum fp32   {t(lambda: C.npm_relu_bwd_colsum(P(h), P(dh), P(dz), P(db), M, F, P(ws), st)):7.1f} us')
print(f'relu_bwd+colsum planes {t(lambda: load().npm_relu_bwd_colsum_planes(P(hp), P(dh), P(dzp), M * F, P(db), M, F, P(ws), st)):7.1f} us')
dxu = dict(b=P(w1), c=P(dx), m=M, n=D, k=F, a_rs=F, a_cs=1, b_rs=1, b_cs=F, ldc=D, flags=0, b_split=P(w1p), b_split_plane=w1.numel())
print(f'dX1 (dz W1^T) A fp32   {t(lambda: gemm(a=P(dz), **dxu)):7.1f} us')
print(f'dX1 (dz W1^T) A planes {t(lambda: gemm(a=None, a_split=P(dzp), a_split_plane=M * F, **dxu)):7.1f} us')
dwu = dict(a=P(x), c=P(dw1), m=D, n=F, k=M, a_rs=1, a_cs=D, b_rs=F, b_cs=1, ldc=F, flags=0)
print(f'dW1 (x^T dz) B fp32   {t(lambda: gemm(b=P(dz), **dwu)):7.1f} us')
print(f'dW1 (x^T dz) B planes {t(lambda: gemm(b=None, b_split=P(dzp), b_split_plane=M * F, **dwu)):7.1f} us')
