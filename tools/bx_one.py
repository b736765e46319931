"""tools/bx_one.py <majors> M N K [prec] — one GEMM in a split-bf16 mode (NPM_GEMM_DEBUG_TIMES=1 prints role wait counters)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tools'))
import bx_check  # noqa: E402
mj, M, N, K = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
prec = sys.argv[5] if len(sys.argv) > 5 else 'bf16x3'
bx_check.run(mj, prec, M, N, K, timing=not os.environ.get('NPM_GEMM_DEBUG_TIMES'))
