"""tools/cublas_probe.py — run cuBLAS TF32 GEMMs at the cfg5 shapes (for an ncu capture of the library's kernel choice:
tile, cluster, shared memory), next to this repo's GEMM.  Diagnostic only; nothing in the product path calls cuBLAS."""
import torch
torch.backends.cuda.matmul.allow_tf32 = True
shapes = [(8192, 4096, 1024), (8192, 1024, 4096), (8192, 1024, 1024), (1024, 4096, 8192)]
for (m, n, k) in shapes:
    a = torch.randn(m, k, device='cuda')
    b = torch.randn(k, n, device='cuda')
    for _ in range(3):
        c = a @ b
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    ts = []
    for _ in range(10):
        flush.zero_()
        e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    print(f'cublas tf32 M={m} N={n} K={k}: {ts[len(ts)//2]*1e3:.1f} us  {2*m*n*k/ts[len(ts)//2]/1e9:.1f} TF', flush=True)
