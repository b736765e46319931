"""tools/ew_bench.py — achieved HBM GB/s of the bandwidth kernels at the cfg5 sizes (8192 rows), CUDA-event timed,
L2 flushed between launches.  Algorithmic bytes per kernel as in DESIGN.md."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'np-modeling_b200'))
from npm_b200._lib import C, TensorEntry, OPT_CHUNK  # noqa: E402

flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
st = torch.cuda.current_stream().cuda_stream
PEAK = 6543.7


ONCE = bool(os.environ.get('NPM_EW_ONCE'))      # under ncu: one launch per kernel, no timing loop


def t(fn, iters=8):
    if ONCE:
        flush.zero_(); fn(); torch.cuda.synchronize()
        return 1.0
    fn(); fn(); ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort(); return ts[len(ts) // 2]


def report(name, nbytes, fn):
    ms = t(fn)
    gbs = nbytes / ms / 1e6
    print(f'{name:46s} {ms * 1e3:8.1f} us  {gbs:8.1f} GB/s  {gbs / PEAK:5.2f} of measured HBM peak', flush=True)


def main():
    R, Cn, F = 8192, 1024, 4096
    n = R * Cn
    x, y, z, w = (torch.randn(R, Cn, device='cuda') for _ in range(4))
    g, b = torch.randn(Cn, device='cuda'), torch.randn(Cn, device='cuda')
    mean, rstd = torch.empty(R, device='cuda'), torch.empty(R, device='cuda')
    dg, db = torch.empty(Cn, device='cuda'), torch.empty(Cn, device='cuda')
    ws = torch.empty(max(C.npm_layernorm_bwd_workspace(R, Cn), C.npm_colsum_workspace(R, F), 16), dtype=torch.uint8, device='cuda')
    mb = torch.empty(C.npm_dropout_layernorm_mask_bytes(R, Cn), dtype=torch.uint8, device='cuda')
    p = lambda a: a.data_ptr()
    report('layernorm_fwd            (8 B/elem)', 8 * n, lambda: C.npm_layernorm_fwd(p(x), p(g), p(b), p(y), p(mean), p(rstd), R, Cn, 1e-3, st))
    report('layernorm_bwd           (12 B/elem)', 12 * n, lambda: C.npm_layernorm_bwd(p(z), p(x), p(g), p(mean), p(rstd), p(y), p(dg), p(db), R, Cn, p(ws), st))
    report('dropout_fwd              (8 B/elem)', 8 * n, lambda: C.npm_dropout_fwd(p(x), p(y), n, 0.9, 1, 0, None, st))
    report('dropout+layernorm fwd    (8 B/elem)', 8 * n, lambda: C.npm_dropout_layernorm_fwd(p(x), p(g), p(b), p(y), p(mean), p(rstd), p(mb), R, Cn, 1e-3, 0.9, 1, 0, st))
    report('dropout+layernorm bwd + residual (16 B/elem)', 16 * n, lambda: C.npm_dropout_layernorm_bwd(p(z), p(x), p(g), p(mean), p(rstd), p(mb), p(w), p(y), p(dg), p(db), R, Cn, 0.9, p(ws), st))
    report('add_inplace             (12 B/elem)', 12 * n, lambda: C.npm_add_inplace(p(y), p(x), n, st))
    report('add3                    (16 B/elem)', 16 * n, lambda: C.npm_add3(p(x), p(z), p(w), p(y), n, st))
    report('colsum [8192,1024]       (4 B/elem)', 4 * n, lambda: C.npm_colsum(p(x), p(dg), R, Cn, p(ws), st))
    X, Y, Z = (torch.randn(R, F, device='cuda') for _ in range(3))
    dbf = torch.empty(F, device='cuda')
    report('colsum [8192,4096]       (4 B/elem)', 4 * R * F, lambda: C.npm_colsum(p(X), p(dbf), R, F, p(ws), st))
    report('relu_bwd+colsum [8192,4096] (12 B/elem)', 12 * R * F, lambda: C.npm_relu_bwd_colsum(p(Y), p(X), p(Z), p(dbf), R, F, p(ws), st))
    sm = torch.randn(R, Cn, device='cuda')
    report('softmax_fwd [8192,1024]  (8 B/elem)', 8 * n, lambda: C.npm_softmax_fwd(p(sm), p(y), R, Cn, st))
    # Adam over one decoder layer's parameters (16.8 M) in one launch: 28 B/param
    sizes = [1024 * 1024] * 8 + [1024 * 4096] * 2 + [4096, 1024] + [1024] * 14
    arr = (TensorEntry * len(sizes))()
    keep = []
    chunks = 0
    for i, s_ in enumerate(sizes):
        ts_ = [torch.randn(s_, device='cuda') for _ in range(2)] + [torch.zeros(s_, device='cuda') for _ in range(2)]
        keep.append(ts_)
        arr[i].param, arr[i].grad, arr[i].m, arr[i].v = (a.data_ptr() for a in ts_)
        arr[i].numel, arr[i].chunk_begin = s_, chunks
        chunks += (s_ + OPT_CHUNK - 1) // OPT_CHUNK
    table = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).cuda()
    report('adam_multi (16.8 M params, 28 B/param)', 28 * sum(sizes), lambda: C.npm_adam_multi(p(table), len(sizes), chunks, 1e-4, 0.9, 0.999, 1e-7, 1, 1.0, st))
    # the same update also rewriting the weights' split-bf16 image (npm_tensor_entry.planes): 32 B/param for the weights
    wsizes = [s_ for s_ in sizes if s_ >= 1024 * 1024]
    planes = [torch.empty(4 * s_, dtype=torch.uint8, device='cuda') for s_ in sizes]
    for i, s_ in enumerate(sizes):
        if s_ >= 1024 * 1024:
            arr[i].planes, arr[i].plane_stride = planes[i].data_ptr(), s_
    table2 = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).cuda()
    report('adam_multi + weight planes (32 B/param)', 28 * sum(sizes) + 4 * sum(wsizes),
           lambda: C.npm_adam_multi(p(table2), len(sizes), chunks, 1e-4, 0.9, 0.999, 1e-7, 1, 1.0, st))
    # split-bf16 activation planes (bf16x3 mode): ReLU backward gated by the hi plane, LayerNorm forward writing planes
    from npm_b200._lib import load
    yp = torch.empty(4 * R * F, dtype=torch.uint8, device='cuda')
    C.npm_weight_split(p(Y), p(yp), R, F, st)
    dzp = torch.empty(4 * R * F, dtype=torch.uint8, device='cuda')
    report('relu_bwd+colsum planes [8192,4096] (10 B/elem)', 10 * R * F,
           lambda: load().npm_relu_bwd_colsum_planes(p(yp), p(X), p(dzp), R * F, p(dbf), R, F, p(ws), st))
    op = torch.empty(4 * n, dtype=torch.uint8, device='cuda')
    report('dropout+layernorm fwd -> planes (8 B/elem)', 8 * n,
           lambda: load().npm_layernorm_fwd_planes(p(x), p(g), p(b), p(op), n, p(mean), p(rstd), p(mb), R, Cn, 1e-3, 0.9, 1, 0, st))


if __name__ == '__main__':
    main()
