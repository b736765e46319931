"""npm_b200 — runtime of the B200-native np-modeling hot path: the ctypes binding of
libnpm_b200.so (`_lib`), device buffers (`device`) and the data-parallel helpers (`dist`)."""
from . import _lib, device  # noqa: F401
from ._lib import NpmError, PREC_3XTF32, PREC_FP32, PREC_TF32  # noqa: F401


def set_precision(mode: str) -> None:
    """'tf32' | '3xtf32' | 'fp32' — contraction precision of the tensor-core paths."""
    table = {'tf32': PREC_TF32, '3xtf32': PREC_3XTF32, 'fp32': PREC_FP32}
    _lib.load().npm_set_precision(table[mode])


def get_precision() -> str:
    return {PREC_TF32: 'tf32', PREC_3XTF32: '3xtf32', PREC_FP32: 'fp32'}[_lib.load().npm_get_precision()]


def launch_count() -> int:
    return int(_lib.load().npm_launch_count())


def reset_launch_count() -> None:
    _lib.load().npm_reset_launch_count()
