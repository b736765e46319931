"""npm_b200 — runtime of the B200-native np-modeling hot path: the ctypes binding of
libnpm_b200.so (`_lib`), device buffers (`device`) and the data-parallel helpers (`dist`)."""
from . import _lib, device  # noqa: F401
from ._lib import NpmError, PREC_3XTF32, PREC_BF16, PREC_BF16X3, PREC_FP32, PREC_TF32  # noqa: F401

_MODES = {'tf32': PREC_TF32, '3xtf32': PREC_3XTF32, 'fp32': PREC_FP32, 'bf16x3': PREC_BF16X3, 'bf16': PREC_BF16}


def set_precision(mode: str) -> None:
    """'bf16x3' | '3xtf32' | 'tf32' | 'bf16' | 'fp32' — contraction precision of the tensor-core paths.
    'bf16x3' and '3xtf32' meet north_star's rtol 1e-3 / atol 1e-4; 'tf32' and 'bf16' are single-pass throughput modes
    with looser stated tolerances (tests/test_fullsize_gpu.py)."""
    _lib.load().npm_set_precision(_MODES[mode])


def get_precision() -> str:
    return {v: k for k, v in _MODES.items()}[_lib.load().npm_get_precision()]


def launch_count() -> int:
    return int(_lib.load().npm_launch_count())


def reset_launch_count() -> None:
    _lib.load().npm_reset_launch_count()
