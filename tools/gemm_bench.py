"""tools/gemm_bench.py — TFLOP/s of the tcgen05 GEMM at the cfg5 shapes (run on the GPU box).
Also measures cuBLAS TF32 (torch.matmul, allow_tf32) as the practical TF32 peak on this box."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'np-modeling_b200'))
from npm_b200._lib import C, GemmDesc  # noqa: E402

PREC = {'tf32': 0, '3xtf32': 1}
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')


def time_fn(fn, iters=8):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def gemm(majors, prec, M, N, K, nb=1):
    a = torch.randn(nb, M, K, device='cuda') if majors[0] == 'k' else torch.randn(nb, K, M, device='cuda')
    b = torch.randn(nb, N, K, device='cuda') if majors[1] == 'k' else torch.randn(nb, K, N, device='cuda')
    c = torch.empty(nb, M, N, device='cuda')
    d = GemmDesc()
    d.a, d.b, d.c, d.bias = a.data_ptr(), b.data_ptr(), c.data_ptr(), None
    d.m, d.n, d.k = M, N, K
    d.a_rs, d.a_cs = (K, 1) if majors[0] == 'k' else (1, M)
    d.b_rs, d.b_cs = (1, K) if majors[1] == 'k' else (N, 1)
    d.ldc = N
    d.nb1, d.nb2 = nb, 1
    d.a_bs1, d.b_bs1, d.c_bs1 = M * K, K * N, M * N
    d.alpha, d.flags, d.precision = 1.0, 0, PREC[prec]
    st = torch.cuda.current_stream().cuda_stream
    ms = time_fn(lambda: C.npm_gemm(d, st))
    return 2.0 * nb * M * N * K / ms / 1e9, ms


def main():
    torch.backends.cuda.matmul.allow_tf32 = True
    for n in (8192, 4096):
        x = torch.randn(n, n, device='cuda'); y = torch.randn(n, n, device='cuda')
        ms = time_fn(lambda: torch.matmul(x, y))
        print(f'cuBLAS tf32 {n}^3: {2.0 * n ** 3 / ms / 1e9:8.1f} TFLOP/s ({ms:.3f} ms)', flush=True)
    xb, yb = torch.randn(8192, 8192, device='cuda', dtype=torch.bfloat16), torch.randn(8192, 8192, device='cuda', dtype=torch.bfloat16)
    ms = time_fn(lambda: torch.matmul(xb, yb))
    print(f'cuBLAS bf16 8192^3: {2.0 * 8192 ** 3 / ms / 1e9:8.1f} TFLOP/s', flush=True)
    cases = [('8192^3', 'kk', 8192, 8192, 8192, 1),
             ('ffn up   fwd  x@W1', 'km', 8192, 4096, 1024, 1), ('ffn down fwd  h@W2', 'km', 8192, 1024, 4096, 1),
             ('ffn up   dX   dy@W1^T', 'kk', 8192, 1024, 4096, 1), ('ffn up   dW   x^T@dy', 'mm', 1024, 4096, 8192, 1),
             ('proj fwd x@Wq^T', 'kk', 8192, 1024, 1024, 1), ('proj dX  dq@Wq', 'km', 8192, 1024, 1024, 1),
             ('proj dW  dq^T@x', 'mm', 1024, 1024, 8192, 1),
             ('attn S=QK^T  (B8 H16)', 'kk', 1024, 1024, 64, 128), ('attn O=PV', 'km', 1024, 64, 1024, 128),
             ('attn dV=P^TdO', 'mm', 1024, 64, 1024, 128), ('attn dQ=dS K', 'km', 1024, 64, 1024, 128)]
    for name, mj, M, N, K, nb in cases:
        row = f'{name:26s} {mj} M={M:5d} N={N:5d} K={K:5d} nb={nb:3d} :'
        if nb == 1:      # cuBLAS TF32 on the same shape / operand majors: the practical ceiling
            a_ = torch.randn(M, K, device='cuda') if mj[0] == 'k' else torch.randn(K, M, device='cuda').t()
            b_ = torch.randn(N, K, device='cuda').t() if mj[1] == 'k' else torch.randn(K, N, device='cuda')
            ms = time_fn(lambda: torch.matmul(a_, b_))
            row += f'  cublas: {2.0 * M * N * K / ms / 1e9:7.1f} TF ({ms:.3f} ms)'
        for prec in ('tf32', '3xtf32'):
            for bn in (0,):
                if bn:
                    os.environ['NPM_GEMM_BLOCK_N_DYN'] = str(bn)
                else:
                    os.environ.pop('NPM_GEMM_BLOCK_N_DYN', None)
                tf, ms = gemm(mj, prec, M, N, K, nb)
                row += f'  {prec}{"/bn" + str(bn) if bn else ""}: {tf:7.1f} TF ({ms:.3f} ms)'
        print(row, flush=True)


if __name__ == '__main__':
    main()
