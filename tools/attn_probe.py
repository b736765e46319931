"""tools/attn_probe.py — attention-core forward/backward against an fp64 torch reference, plus timing
(run on the GPU box).  usage: python tools/attn_probe.py [B H Sq Skv] [--bwd] [--time]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'np-modeling_b200'))
import npm_b200  # noqa: E402
from npm_b200._lib import C  # noqa: E402

D = 64


def ref_fwd(q, k, v):
    qd, kd, vd = q.double(), k.double(), v.double()
    s = torch.einsum('bshd,bthd->bhst', qd, kd) / D ** 0.5
    p = torch.softmax(s, dim=-1)
    return torch.einsum('bhst,bthd->bshd', p, vd), p


def stats(name, got, want):
    err = (got.double() - want).abs()
    print(f'  {name:4s} max_abs_err={err.max().item():.3e}  rms_err={err.pow(2).mean().sqrt().item():.3e}  '
          f'ref_rms={want.pow(2).mean().sqrt().item():.3e}  nan={int(torch.isnan(got).sum())}', flush=True)
    return err.max().item()


def run(B, H, Sq, Skv, bwd=False, timing=False, seed=0):
    g = torch.Generator(device='cuda').manual_seed(seed)
    q = torch.randn(B, Sq, H, D, generator=g, device='cuda')
    k = torch.randn(B, Skv, H, D, generator=g, device='cuda')
    v = torch.randn(B, Skv, H, D, generator=g, device='cuda')
    do = torch.randn(B, Sq, H, D, generator=g, device='cuda')
    if '--bias' in sys.argv:      # a large component common to all positions (what projection biases produce)
        q = q + 0.7 * torch.randn(1, 1, H, D, generator=g, device='cuda')
        k = k + 0.7 * torch.randn(1, 1, H, D, generator=g, device='cuda')
        v = v + 0.7 * torch.randn(1, 1, H, D, generator=g, device='cuda')
        q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
    o = torch.full((B, Sq, H, D), float('nan'), device='cuda')
    st = torch.cuda.current_stream().cuda_stream
    saved = torch.empty(C.npm_mha_core_saved_bytes(B, H, Sq, Skv, D, D), dtype=torch.uint8, device='cuda')
    C.npm_mha_core_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), saved.data_ptr(), B, H, Sq, Skv, D, D, st)
    torch.cuda.synchronize()
    print(f'B={B} H={H} Sq={Sq} Skv={Skv}  fused={os.environ.get("NPM_ATTN_FUSED", "0")}', flush=True)
    small = B * H * Sq * Skv <= (1 << 28)
    if small:
        ro, rp = ref_fwd(q, k, v)
        stats('o', o, ro)
    if bwd:
        dq = torch.full_like(q, float('nan')); dk = torch.full_like(k, float('nan')); dv = torch.full_like(v, float('nan'))
        scratch = torch.empty(C.npm_mha_core_bwd_scratch_bytes(B, H, Sq, Skv, D, D), dtype=torch.uint8, device='cuda')
        C.npm_mha_core_bwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), do.data_ptr(), saved.data_ptr(),
                           dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), scratch.data_ptr(), B, H, Sq, Skv, D, D, st)
        torch.cuda.synchronize()
        if small:
            dod = do.double()
            rdv = torch.einsum('bhst,bshd->bthd', rp, dod)
            dp = torch.einsum('bshd,bthd->bhst', dod, v.double())
            ds = rp * (dp - (dp * rp).sum(-1, keepdim=True)) / D ** 0.5
            rdq = torch.einsum('bhst,bthd->bshd', ds, k.double())
            rdk = torch.einsum('bhst,bshd->bthd', ds, q.double())
            stats('dq', dq, rdq); stats('dk', dk, rdk); stats('dv', dv, rdv)
    if timing:
        flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')

        def t(fn, iters=10):
            fn(); ts = []
            for _ in range(iters):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ts.sort(); return ts[len(ts) // 2]
        ms = t(lambda: C.npm_mha_core_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), saved.data_ptr(),
                                          B, H, Sq, Skv, D, D, st))
        fl = 4.0 * B * H * Sq * Skv * D
        print(f'  fwd {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s', flush=True)
        if '--causal' in sys.argv and Sq == Skv:
            import ctypes
            from npm_b200._lib import MhaStrides
            ld = MhaStrides(causal=1)
            msc = t(lambda: C.npm_mha_core_fwd_strided(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), saved.data_ptr(),
                                                       B, H, Sq, Skv, D, D, ctypes.byref(ld), st))
            print(f'  causal fwd {msc:.3f} ms  ({ms / msc:.2f}x the unmasked launch)', flush=True)
            if bwd:
                msb = t(lambda: C.npm_mha_core_bwd_strided(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), do.data_ptr(),
                                                           saved.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(),
                                                           scratch.data_ptr(), B, H, Sq, Skv, D, D, ctypes.byref(ld), st))
                print(f'  causal bwd {msb:.3f} ms', flush=True)
        if bwd:
            ms = t(lambda: C.npm_mha_core_bwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), do.data_ptr(),
                                              saved.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(),
                                              scratch.data_ptr(), B, H, Sq, Skv, D, D, st))
            print(f'  bwd {ms:.3f} ms  {2.5 * fl / ms / 1e9:.1f} TFLOP/s (5 GEMM units)', flush=True)


if __name__ == '__main__':
    prec = [a.split('=')[1] for a in sys.argv if a.startswith('--prec=')]
    npm_b200.set_precision(prec[0] if prec else 'tf32')
    args = [a for a in sys.argv[1:] if not a.startswith('--')]
    dims = [int(a) for a in args] if args else [2, 4, 256, 384]
    run(*dims, bwd='--bwd' in sys.argv, timing='--time' in sys.argv)
