"""tools/bx_check.py — numerics (vs fp64) and TFLOP/s of the split-bf16 GEMM (gemm_bx.cu) for every operand-major
combination, ragged edges, batches, split-K, bias / residual / ReLU epilogues.  Run on the GPU box."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'np-modeling_b200'))
from npm_b200._lib import C, GemmDesc  # noqa: E402

PREC = {'tf32': 0, '3xtf32': 1, 'fp32': 2, 'bf16x3': 3, 'bf16': 4}
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')


def time_fn(fn, iters=6):
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


def run(majors, prec, M, N, K, nb=1, bias=False, residual=False, relu=False, timing=False, seed=0, presplit=False, presplit_a=False):
    g = torch.Generator(device='cuda').manual_seed(seed)
    a = torch.randn(nb, M, K, device='cuda', generator=g) if majors[0] == 'k' else torch.randn(nb, K, M, device='cuda', generator=g)
    b = torch.randn(nb, N, K, device='cuda', generator=g) if majors[1] == 'k' else torch.randn(nb, K, N, device='cuda', generator=g)
    c = torch.full((nb, M, N), float('nan'), device='cuda')
    bv = torch.randn(N, device='cuda', generator=g) if bias else None
    rv = torch.randn(M, N, device='cuda', generator=g) if residual else None
    d = GemmDesc()
    d.a, d.b, d.c = a.data_ptr(), b.data_ptr(), c.data_ptr()
    d.bias = bv.data_ptr() if bias else None
    d.m, d.n, d.k = M, N, K
    d.a_rs, d.a_cs = (K, 1) if majors[0] == 'k' else (1, M)
    d.b_rs, d.b_cs = (1, K) if majors[1] == 'k' else (N, 1)
    d.ldc = N
    d.nb1, d.nb2 = nb, 1
    d.a_bs1, d.b_bs1, d.c_bs1 = M * K, K * N, M * N
    d.alpha, d.flags, d.precision = 1.0, (1 if relu else 0), PREC[prec]
    d.residual = rv.data_ptr() if residual else None
    d.ldr = N
    if presplit:      # B's bf16 hi / mid planes (npm_weight_split), as the layers pass them for weights
        planes = torch.empty(b.numel() * 4, dtype=torch.uint8, device='cuda')
        C.npm_weight_split(b.data_ptr(), planes.data_ptr(), b.shape[-2], b.shape[-1], torch.cuda.current_stream().cuda_stream)
        d.b_split, d.b_split_plane = planes.data_ptr(), b.numel()
    if presplit_a:    # A as bf16 planes, no fp32 pointer at all (what a producer that writes planes hands over)
        planes_a = torch.empty(a.numel() * 4, dtype=torch.uint8, device='cuda')
        C.npm_weight_split(a.data_ptr(), planes_a.data_ptr(), a.shape[-2], a.shape[-1], torch.cuda.current_stream().cuda_stream)
        d.a_split, d.a_split_plane, d.a = planes_a.data_ptr(), a.numel(), None
    st = torch.cuda.current_stream().cuda_stream
    C.npm_gemm(d, st)
    torch.cuda.synchronize()
    A = a.double() if majors[0] == 'k' else a.double().transpose(1, 2)
    B = b.double().transpose(1, 2) if majors[1] == 'k' else b.double()
    want = A @ B
    if bias:
        want = want + bv.double()
    if residual:
        want = want + rv.double()
    if relu:
        want = want.clamp_min(0)
    got = c.double()
    err = (got - want).abs()
    tol = 1e-4 + 1e-3 * want.abs()
    viol = float((err / tol).max())
    rel = float(err.max() / want.abs().max())
    tag = prec + ('+A' if presplit_a else '') + ('+B' if presplit else '')
    out = f'{tag:11s} {majors} M={M:5d} N={N:5d} K={K:5d} nb={nb:3d} b{int(bias)}r{int(residual)}relu{int(relu)}: max|err|={float(err.max()):.3e} ' \
          f'rel-to-max={rel:.2e} worst err/tol={viol:.3f} nan={int(torch.isnan(c).sum())}'
    if timing:
        ms = time_fn(lambda: C.npm_gemm(d, st))
        out += f'  {2.0 * nb * M * N * K / ms / 1e9:7.1f} TF ({ms:.3f} ms)'
    print(out, flush=True)
    return viol


def main():
    bad = 0
    for mj, M, N, K in [('kk', 256, 256, 64), ('km', 256, 256, 64), ('kk', 300, 520, 1024), ('km', 300, 512, 104), ('km', 1000, 1024, 1000), ('kk', 513, 36, 96)]:
        bad += run(mj, 'bf16x3', M, N, K, presplit=True, bias=True) > 1.0 and M * K < 100000
    for mj, M, N, K in [('kk', 256, 256, 64), ('km', 256, 256, 64), ('mk', 256, 256, 64), ('mm', 256, 256, 64), ('kk', 300, 520, 1024),
                        ('mm', 1024, 1024, 1000), ('mk', 512, 104, 264), ('km', 1000, 1024, 1000)]:
        for pa, pb in ((True, False), (True, True)):
            bad += run(mj, 'bf16x3', M, N, K, presplit=pb, presplit_a=pa, bias=True) > 1.0 and K < 200
    quick = [('kk', 256, 256, 64), ('km', 256, 256, 64), ('mk', 256, 256, 64), ('mm', 256, 256, 64),
             ('kk', 300, 520, 1024), ('km', 300, 520, 100), ('mk', 260, 36, 516), ('mm', 1024, 1024, 1024),
             ('kk', 129, 8, 4), ('km', 1000, 1000, 1000), ('mm', 516, 260, 132)]
    for prec in ('bf16x3', 'bf16'):
        for mj, M, N, K in quick:
            v = run(mj, prec, M, N, K)
            bad += (prec == 'bf16x3' and v > 1.0)
    for mj in ('kk', 'km', 'mk', 'mm'):
        bad += run(mj, 'bf16x3', 384, 320, 96, nb=5) > 1.0
    bad += run('kk', 'bf16x3', 512, 512, 256, bias=True, residual=True) > 1.0
    bad += run('km', 'bf16x3', 512, 512, 256, bias=True, relu=True) > 1.0
    bad += run('mm', 'bf16x3', 1024, 1024, 8192) > 1.0          # split-K
    print('--- timing (cfg5 shapes) ---', flush=True)
    cases = [('kk', 8192, 8192, 8192), ('km', 8192, 4096, 1024), ('km', 8192, 1024, 4096), ('kk', 8192, 1024, 4096),
             ('mm', 1024, 4096, 8192), ('kk', 8192, 1024, 1024), ('kk', 8192, 3072, 1024), ('km', 8192, 1024, 1024), ('mm', 1024, 1024, 8192)]
    for mj, M, N, K in cases:
        for prec in ('bf16x3', 'bf16', 'tf32', '3xtf32'):
            run(mj, prec, M, N, K, timing=True)
        run(mj, 'bf16x3', M, N, K, timing=True, presplit=True)
        run(mj, 'bf16x3', M, N, K, timing=True, presplit_a=True)
        run(mj, 'bf16x3', M, N, K, timing=True, presplit=True, presplit_a=True)
    print('BX_CHECK', 'FAIL' if bad else 'OK', bad, flush=True)


if __name__ == '__main__':
    main()
