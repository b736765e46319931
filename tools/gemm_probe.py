"""tools/gemm_probe.py — diagnostic for the tcgen05 GEMM (run on the GPU box).

usage: python tools/gemm_probe.py CASE   where CASE = "<a_major><b_major>:<prec>:<M>x<N>x<K>[:batch]"
  majors: k = contraction index contiguous (K-major), m = MN-major.   prec: tf32 | 3xtf32 | fp32
Prints max-abs / relative error against an fp64 torch matmul and, on failure, decodes where
one-hot inputs land (to debug descriptor / swizzle layouts).
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'np-modeling_b200'))
from npm_b200 import _lib  # noqa: E402
from npm_b200._lib import C, GemmDesc  # noqa: E402

PREC = {'tf32': 0, '3xtf32': 1, 'fp32': 2}


def run(a_major, b_major, prec, M, N, K, nb=1, alpha=1.0, bias=False, relu=False, accum=False, seed=0, onehot=None):
    g = torch.Generator(device='cuda').manual_seed(seed)
    # logical A [nb, M, K], B [nb, K, N]
    if onehot is None:
        A = torch.randn(nb, M, K, generator=g, device='cuda')
        B = torch.randn(nb, K, N, generator=g, device='cuda')
    else:
        m0, k0 = onehot
        A = torch.zeros(nb, M, K, device='cuda')
        A[:, m0, k0] = 1.0
        B = (torch.arange(K, device='cuda').float()[:, None] * 1000 + torch.arange(N, device='cuda').float()[None, :])
        B = B[None].repeat(nb, 1, 1).contiguous()
    a_store = A.contiguous() if a_major == 'k' else A.transpose(1, 2).contiguous()      # [nb,M,K] or [nb,K,M]
    b_store = B.transpose(1, 2).contiguous() if b_major == 'k' else B.contiguous()      # [nb,N,K] or [nb,K,N]
    Cout = torch.full((nb, M, N), 7.0 if accum else float('nan'), device='cuda')
    bias_t = torch.randn(N, generator=g, device='cuda') if bias else None
    d = GemmDesc()
    d.a, d.b, d.c = a_store.data_ptr(), b_store.data_ptr(), Cout.data_ptr()
    d.bias = bias_t.data_ptr() if bias else None
    d.m, d.n, d.k = M, N, K
    d.a_rs, d.a_cs = (K, 1) if a_major == 'k' else (1, M)
    d.b_rs, d.b_cs = (1, K) if b_major == 'k' else (N, 1)
    d.ldc = N
    d.nb1, d.nb2 = nb, 1
    d.a_bs1, d.b_bs1, d.c_bs1 = M * K, K * N, M * N
    d.alpha = alpha
    d.flags = (1 if relu else 0) | (2 if accum else 0)
    d.precision = PREC[prec]
    C.npm_gemm(d, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    ref = alpha * torch.matmul(A.double(), B.double())
    if bias:
        ref = ref + bias_t.double()
    if relu:
        ref = ref.clamp_min(0)
    if accum:
        ref = ref + 7.0
    return Cout, ref


def one(case):
    parts = case.split(':')
    majors, prec, dims = parts[0], parts[1], parts[2]
    nb = int(parts[3]) if len(parts) > 3 else 1
    M, N, K = (int(v) for v in dims.split('x'))
    extras = set(parts[4].split(',')) if len(parts) > 4 else set()
    out, ref = run(majors[0], majors[1], prec, M, N, K, nb, alpha=0.5 if 'alpha' in extras else 1.0,
                   bias='bias' in extras, relu='relu' in extras, accum='accum' in extras)
    err = (out.double() - ref).abs()
    scale = ref.abs().max().item() + 1e-30
    nan = int(torch.isnan(out).sum().item())
    print(f'CASE {case:40s} max_abs_err={err.max().item():.3e} rel_to_max={err.max().item() / scale:.3e} '
          f'rms_err={err.pow(2).mean().sqrt().item():.3e} ref_rms={ref.pow(2).mean().sqrt().item():.3e} nan={nan}',
          flush=True)
    tol = {'tf32': 4e-3, '3xtf32': 2e-5, 'fp32': 2e-5}[prec] * (K ** 0.5) * 4
    if nan or err.max().item() > tol:
        print('  FAIL — decoding with one-hot A(m0=5,k0=3), B[k,n] = 1000k + n', flush=True)
        o2, r2 = run(majors[0], majors[1], prec, M, N, K, nb, onehot=(5, 3))
        row = o2[0, 5, :12].tolist()
        print('   got  C[5,:12] =', [round(v, 1) for v in row])
        print('   want C[5,:12] =', [round(v, 1) for v in r2[0, 5, :12].tolist()])
        nz = torch.nonzero(o2[0].nan_to_num(0) != 0)
        print('   nonzero rows:', sorted(set(nz[:, 0].tolist()))[:16], ' count', nz.shape[0])
        for k0 in (0, 8, 9, 31, 33):
            if k0 < K:
                o3, _ = run(majors[0], majors[1], prec, M, N, K, nb, onehot=(70 % M, k0))
                print(f'   onehot(m0={70 % M},k0={k0}): C[m0,:4] =', [round(v, 1) for v in o3[0, 70 % M, :4].tolist()],
                      ' C[m0,40:44] =', [round(v, 1) for v in o3[0, 70 % M, 40:44].tolist()] if N > 44 else '')
        return False
    return True


CASES = [f'{mj}:{prec}:{dims}' for prec in ('tf32', '3xtf32') for mj in ('kk', 'km', 'mk', 'mm')
         for dims in ('128x64x32', '128x128x64', '128x256x96', '256x512x256', '1000x520x264')] + [
    'kk:tf32:512x512x512:3', 'km:3xtf32:300x200x100:5', 'kk:3xtf32:2048x2048x2048', 'kk:tf32:4096x4096x4096',
    'kk:3xtf32:256x256x128:1:bias,relu,alpha', 'kk:tf32:256x256x128:2:accum', 'mm:3xtf32:256x256x128:2:accum,bias',
    'kk:fp32:100x30x50', 'km:fp32:64x10x256', 'kk:3xtf32:64x12x256', 'mk:3xtf32:8192x1024x4096', 'km:tf32:8192x4096x1024']


def main():
    """`--all`: parent restarts a worker after any crash so one trapped kernel costs one case."""
    import subprocess
    if sys.argv[1] == '--all':
        i = 0
        while i < len(CASES):
            p = subprocess.Popen([sys.executable, __file__, '--worker', str(i)], stdout=subprocess.PIPE, text=True)
            for line in p.stdout:
                print(line, end='', flush=True)
                if line.startswith('DONE '):
                    i = int(line.split()[1]) + 1
            p.wait()
            if p.returncode != 0 and i < len(CASES):
                print(f'  WORKER DIED (rc={p.returncode}) on case {CASES[i]}', flush=True)
                i += 1
    elif sys.argv[1] == '--worker':
        for i in range(int(sys.argv[2]), len(CASES)):
            try:
                one(CASES[i])
            except _lib.NpmError as e:
                print(f'CASE {CASES[i]} ERROR {e}', flush=True)
                if 'launch failure' in str(e) or 'illegal' in str(e):
                    sys.exit(3)
            print(f'DONE {i}', flush=True)
    else:
        sys.exit(0 if one(sys.argv[1]) else 1)


if __name__ == '__main__':
    main()
