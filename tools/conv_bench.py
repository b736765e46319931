"""tools/conv_bench.py — Conv2D fprop / dgrad / wgrad TFLOP/s at the BASELINE cfg2 shapes (TF32 tensor-core path vs
the fp32 CUDA-core kernel), CUDA-event timed, L2 flushed."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'np-modeling_b200'))
import npm_b200  # noqa: E402
from npm_b200._lib import C  # noqa: E402

flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
st = torch.cuda.current_stream().cuda_stream


def t(fn, iters=6):
    fn(); fn(); ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort(); return ts[len(ts) // 2]


def main():
    for (N, H, W, Ci, Co, k) in [(256, 32, 32, 64, 128, 3), (256, 32, 32, 3, 64, 3), (64, 64, 64, 128, 128, 3)]:
        x = torch.randn(N, H, W, Ci, device='cuda'); dy = torch.randn(N, H, W, Co, device='cuda')
        f = torch.randn(k, k, Ci, Co, device='cuda') / (k * k * Ci) ** 0.5
        b = torch.randn(Co, device='cuda')
        y = torch.empty(N, H, W, Co, device='cuda'); dx = torch.empty_like(x); dw = torch.empty_like(f); db = torch.empty(Co, device='cuda')
        ws = torch.empty(max(C.npm_conv2d_workspace(N, H, W, Ci, Co, k), 16), dtype=torch.uint8, device='cuda')
        fl = 2.0 * N * H * W * k * k * Ci * Co
        for mode in ('bf16x3', 'tf32', '3xtf32'):
            npm_b200.set_precision(mode)
            p = lambda a: a.data_ptr()
            a = t(lambda: C.npm_conv2d_fwd(p(x), p(f), p(b), p(y), N, H, W, Ci, Co, k, 0, p(ws), st))
            bb = t(lambda: C.npm_conv2d_bwd_dx(p(dy), p(f), p(dx), N, H, W, Ci, Co, k, p(ws), st))
            c = t(lambda: C.npm_conv2d_bwd_dw_db(p(x), p(dy), p(dw), p(db), N, H, W, Ci, Co, k, p(ws), st))
            path = 'tcgen05 implicit GEMM' if mode in ('tf32', 'bf16x3') and Ci % 4 == 0 else 'fp32 CUDA-core kernel'
            print(f'x[{N},{H},{W},{Ci}] k{k} -> {Co}  {mode:7s} ({path}): fprop {a:7.3f} ms {fl / a / 1e9:6.1f} TF | '
                  f'dgrad {bb:7.3f} ms {fl / bb / 1e9:6.1f} TF | wgrad+db {c:7.3f} ms {fl / c / 1e9:6.1f} TF', flush=True)


if __name__ == '__main__':
    main()
