"""ctypes binding of libnpm_b200.so (C-ABI declared in include/npm_b200.h).

There is deliberately no fallback: if the shared library is missing or a call
fails, an exception is raised — the product path never computes on the CPU.
"""
import ctypes
import os
from ctypes import (POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_uint64,
                    c_void_p)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('NPM_B200_LIB') or os.path.join(os.path.dirname(_HERE), 'libnpm_b200.so')   # env override: tools only

NPM_OK = 0
PREC_TF32, PREC_3XTF32, PREC_FP32, PREC_BF16X3, PREC_BF16 = 0, 1, 2, 3, 4
GEMM_RELU, GEMM_ACCUM = 1, 2
OPT_CHUNK = 8192


class NpmError(RuntimeError):
    pass


class GemmDesc(Structure):
    _fields_ = [
        ('a', c_void_p), ('b', c_void_p), ('c', c_void_p), ('bias', c_void_p),
        ('m', c_int64), ('n', c_int64), ('k', c_int64),
        ('a_rs', c_int64), ('a_cs', c_int64), ('b_rs', c_int64), ('b_cs', c_int64),
        ('ldc', c_int64),
        ('nb1', c_int32), ('nb2', c_int32),
        ('a_bs1', c_int64), ('a_bs2', c_int64), ('b_bs1', c_int64), ('b_bs2', c_int64),
        ('c_bs1', c_int64), ('c_bs2', c_int64),
        ('alpha', c_float), ('flags', c_int32), ('precision', c_int32),
        ('residual', c_void_p), ('ldr', c_int64), ('a_colsum', c_void_p), ('b_split', c_void_p), ('b_split_plane', c_int64),
        ('a_split', c_void_p), ('a_split_plane', c_int64), ('c_split', c_void_p), ('c_split_plane', c_int64),
        ('rowdot_x', c_void_p), ('rowdot_ld', c_int64), ('rowdot_out', c_void_p), ('rowdot_seq', c_int64),
    ]


class MhaStrides(Structure):
    """npm_mha_strides (include/npm_b200.h): floats between consecutive tokens, 0 = dense."""
    _fields_ = [('q', c_int64), ('k', c_int64), ('v', c_int64), ('dq', c_int64), ('dk', c_int64), ('dv', c_int64),
                ('causal', c_int64), ('path', c_int64), ('planes', c_int64), ('q_plane', c_int64), ('k_plane', c_int64),
                ('v_plane', c_int64), ('do_ready', c_int64)]


class TensorEntry(Structure):
    _fields_ = [('param', c_void_p), ('grad', c_void_p), ('m', c_void_p), ('v', c_void_p),
                ('numel', c_int64), ('chunk_begin', c_int64), ('planes', c_void_p), ('plane_stride', c_int64)]


P, I64, F, I, U64, SZ = c_void_p, c_int64, c_float, c_int, c_uint64, c_size_t

# name -> (restype, argtypes).  Every symbol include/npm_b200.h declares is listed here;
# tests/test_abi.py checks the header and this table against the built library.
SIGNATURES = {
    'npm_last_error': (c_char_p, []),
    'npm_version': (c_int, []),
    'npm_launch_count': (c_uint64, []),
    'npm_reset_launch_count': (None, []),
    'npm_set_precision': (c_int, [I]),
    'npm_get_precision': (c_int, []),
    'npm_device_info': (c_int, [POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    'npm_gemm': (c_int, [POINTER(GemmDesc), P]),
    'npm_linear_fwd': (c_int, [P, P, P, P, I64, I64, I64, I, I, P]),
    'npm_linear_fwd_residual': (c_int, [P, P, P, P, P, I64, I64, I64, I, P]),
    'npm_linear_bwd_dx': (c_int, [P, P, P, I64, I64, I64, I, P]),
    'npm_weight_split_bytes': (c_size_t, [I64, I64]),
    'npm_weight_split': (c_int, [P, P, I64, I64, P]),
    'npm_linear_fwd_presplit': (c_int, [P, P, P, I64, P, P, P, I64, I64, I64, I, I, P]),
    'npm_linear_bwd_dx_presplit': (c_int, [P, P, P, I64, P, I64, I64, I64, I, P]),
    'npm_linear_fwd_planes': (c_int, [P, P, P, I64, P, P, I64, I64, I64, I64, I, P]),
    'npm_linear_bwd_dw_db': (c_int, [P, P, P, P, I64, I64, I64, I, P, P]),
    'npm_linear_bwd_dx_planes_rowdot': (c_int, [P, P, P, I64, P, I64, I64, I64, I64, I, P, I64, P, I64, P]),
    'npm_relu_bwd_colsum_planes': (c_int, [P, P, P, I64, P, I64, I64, P, P]),
    'npm_layernorm_fwd_planes': (c_int, [P, P, P, P, I64, P, P, P, I64, I64, F, F, U64, U64, P]),
    'npm_planes_join': (c_int, [P, I64, P, I64, P]),
    'npm_colsum_workspace': (c_size_t, [I64, I64]),
    'npm_colsum': (c_int, [P, P, I64, I64, P, P]),
    'npm_relu_fwd': (c_int, [P, P, I64, P]),
    'npm_relu_bwd': (c_int, [P, P, P, I64, P]),
    'npm_relu_bwd_y': (c_int, [P, P, P, I64, P]),
    'npm_relu_bwd_colsum': (c_int, [P, P, P, P, I64, I64, P, P]),
    'npm_softmax_fwd': (c_int, [P, P, I64, I64, P]),
    'npm_softmax_bwd': (c_int, [P, P, P, I64, I64, F, P]),
    'npm_layernorm_fwd': (c_int, [P, P, P, P, P, P, I64, I64, F, P]),
    'npm_layernorm_bwd_workspace': (c_size_t, [I64, I64]),
    'npm_layernorm_bwd': (c_int, [P, P, P, P, P, P, P, P, I64, I64, P, P]),
    'npm_dropout_layernorm_fused': (c_int, [I64, I64]),
    'npm_dropout_layernorm_mask_bytes': (c_size_t, [I64, I64]),
    'npm_dropout_layernorm_fwd': (c_int, [P, P, P, P, P, P, P, I64, I64, F, F, U64, U64, P]),
    'npm_dropout_layernorm_bwd': (c_int, [P, P, P, P, P, P, P, P, P, P, I64, I64, F, P, P]),
    'npm_dropout_layernorm_bwd_colsum': (c_int, [P, P, P, P, P, P, P, P, P, P, P, I64, I64, F, P, P]),
    'npm_dropout_fwd': (c_int, [P, P, I64, F, U64, U64, P, P]),
    'npm_dropout_bwd': (c_int, [P, P, I64, F, U64, U64, P, P]),
    'npm_dropout_mask': (c_int, [P, I64, F, U64, U64, P]),
    'npm_add_inplace': (c_int, [P, P, I64, P]),
    'npm_add3': (c_int, [P, P, P, P, I64, P]),
    'npm_scale': (c_int, [P, F, I64, P]),
    'npm_fill': (c_int, [P, F, I64, P]),
    'npm_embedding_fwd': (c_int, [P, P, P, I64, I64, I64, P]),
    'npm_embedding_bwd': (c_int, [P, P, P, I64, I64, I64, P]),
    'npm_mha_core_path': (c_int, [I64] * 6),
    'npm_mha_core_saved_bytes_for': (c_size_t, [I] + [I64] * 6),
    'npm_mha_core_bwd_scratch_bytes_for': (c_size_t, [I] + [I64] * 6),
    'npm_mha_core_saved_bytes': (c_size_t, [I64] * 6),
    'npm_mha_core_bwd_scratch_bytes': (c_size_t, [I64] * 6),
    'npm_mha_core_fwd': (c_int, [P, P, P, P, P] + [I64] * 6 + [P]),
    'npm_mha_core_bwd': (c_int, [P] * 10 + [I64] * 6 + [P]),
    'npm_mha_core_fwd_strided': (c_int, [P, P, P, P, P] + [I64] * 6 + [P, P]),
    'npm_mha_core_bwd_strided': (c_int, [P] * 10 + [I64] * 6 + [P, P]),
    'npm_mha_core_scores': (c_int, [P, P, P, P] + [I64] * 6 + [P]),
    'npm_conv2d_workspace': (c_size_t, [I64, I64, I64, I64, I64, I]),
    'npm_conv2d_fwd': (c_int, [P, P, P, P, I64, I64, I64, I64, I64, I, I, P, P]),
    'npm_conv2d_bwd_dx': (c_int, [P, P, P, I64, I64, I64, I64, I64, I, P, P]),
    'npm_conv2d_bwd_dw_db': (c_int, [P, P, P, P, I64, I64, I64, I64, I64, I, P, P]),
    'npm_mse_fwd': (c_int, [P, P, P, I64, P]),
    'npm_mse_bwd': (c_int, [P, P, P, I64, P]),
    'npm_ce_fwd': (c_int, [P, P, P, I64, P]),
    'npm_ce_bwd': (c_int, [P, P, P, I64, P]),
    'npm_sgd_multi': (c_int, [P, c_int32, I64, F, F, P]),
    'npm_adam_multi': (c_int, [P, c_int32, I64, F, F, F, F, c_int32, F, P]),
}

_lib = None


def load():
    """Load the shared library (once). Raises NpmError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NpmError(
            f'{LIB_PATH} is missing: build it with `python -c "import __graft_entry__ as g; g.build()"` '
            f'or `make -C np-modeling_b200/csrc`. There is no CPU fallback.')
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error():
    return load().npm_last_error().decode('utf-8', 'replace')


def check(rc, what=''):
    if rc != NPM_OK:
        raise NpmError(f'{what} failed (rc={rc}): {last_error()}')


class _Calls:
    """`C.npm_xxx(args...)` calls the C function and raises NpmError on a non-zero return code."""

    def __getattr__(self, name):
        lib = load()
        fn = getattr(lib, name)
        res = SIGNATURES[name][0]
        if res is c_int and name not in ('npm_version', 'npm_set_precision', 'npm_get_precision', 'npm_dropout_layernorm_fused', 'npm_mha_core_path', 'npm_linear_fwd_planes', 'npm_linear_bwd_dx_planes_rowdot', 'npm_relu_bwd_colsum_planes', 'npm_layernorm_fwd_planes'):
            def call(*args, _fn=fn, _name=name):
                rc = _fn(*args)
                if rc != NPM_OK:
                    raise NpmError(f'{_name} failed (rc={rc}): {last_error()}')
            wrapped = call
        else:
            wrapped = fn
        setattr(self, name, wrapped)
        return wrapped


C = _Calls()
