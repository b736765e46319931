"""tools/bx_one.py <majors> M N K [prec] — one GEMM in a split-bf16 mode (NPM_GEMM_DEBUG_TIMES=1 prints role wait counters)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tools'))
import bx_check  # noqa: E402
argv = [a for a in sys.argv if not a.startswith('--')]
mj, M, N, K = argv[1], int(argv[2]), int(argv[3]), int(argv[4])
prec = argv[5] if len(argv) > 5 else 'bf16x3'
bx_check.run(mj, prec, M, N, K, timing=not os.environ.get('NPM_GEMM_DEBUG_TIMES'), presplit='--presplit' in sys.argv)
