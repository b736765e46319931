"""The C-ABI library loads without a GPU and exports every symbol include/npm_b200.h declares
(no compute calls here); the ctypes table in npm_b200/_lib.py covers exactly the same set."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'npm_b200.h')


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(npm_[a-z0-9_]+)\s*\(', text)))


def test_header_declares_the_survey_minimum():
    syms = set(declared_symbols())
    for need in ['npm_linear_fwd', 'npm_linear_bwd_dx', 'npm_linear_bwd_dw_db', 'npm_conv2d_fwd', 'npm_conv2d_bwd_dx',
                 'npm_conv2d_bwd_dw_db', 'npm_mha_core_fwd', 'npm_mha_core_bwd', 'npm_layernorm_fwd',
                 'npm_layernorm_bwd', 'npm_softmax_fwd', 'npm_softmax_bwd', 'npm_relu_fwd', 'npm_relu_bwd',
                 'npm_dropout_fwd', 'npm_dropout_bwd', 'npm_mse_fwd', 'npm_mse_bwd', 'npm_ce_fwd', 'npm_ce_bwd',
                 'npm_sgd_multi', 'npm_adam_multi', 'npm_add_inplace', 'npm_last_error']:
        assert need in syms, need


def test_library_exports_every_declared_symbol():
    from npm_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), f'{name} is declared in include/npm_b200.h but not exported'


def test_ctypes_table_matches_header():
    from npm_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()


def test_struct_layouts_match_the_header():
    from npm_b200 import _lib
    assert ctypes.sizeof(_lib.TensorEntry) == 64          # 4 pointers + 2 int64 + planes pointer + plane stride
    assert ctypes.sizeof(_lib.MhaStrides) == 13 * 8       # q k v dq dk dv causal path planes q_plane k_plane v_plane do_ready
    assert _lib.MhaStrides.path.offset == 56
    assert _lib.MhaStrides.causal.offset == 48
    # incl. padding before `residual`; then a_colsum, b_split, b_split_plane
    # incl. a_split, a_split_plane, c_split, c_split_plane
    assert ctypes.sizeof(_lib.GemmDesc) == 4 * 8 + 3 * 8 + 5 * 8 + 2 * 4 + 6 * 8 + 3 * 4 + 4 + 2 * 8 + 7 * 8 + 4 * 8
    assert _lib.GemmDesc.residual.offset == ctypes.sizeof(_lib.GemmDesc) - 16 - 56 - 32
    assert _lib.GemmDesc.a_colsum.offset == ctypes.sizeof(_lib.GemmDesc) - 56 - 32
    assert _lib.GemmDesc.c_split.offset == ctypes.sizeof(_lib.GemmDesc) - 16 - 32
    assert _lib.GemmDesc.alpha.offset == 4 * 8 + 3 * 8 + 5 * 8 + 8 + 6 * 8


def test_state_calls_work_without_a_gpu():
    from npm_b200 import _lib
    lib = _lib.load()
    assert lib.npm_version() >= 100
    prev = lib.npm_set_precision(_lib.PREC_TF32)
    assert lib.npm_get_precision() == _lib.PREC_TF32
    lib.npm_set_precision(prev)
    lib.npm_reset_launch_count()
    assert lib.npm_launch_count() == 0
    assert lib.npm_colsum_workspace(0, 0) == 0


def test_product_path_fails_loudly_without_cuda():
    """No CPU fallback: on a box without a GPU the layers raise instead of computing on the host."""
    import numpy as np
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    from layers import Dense
    with pytest.raises(RuntimeError):
        Dense(4)(np.zeros((2, 3), np.float32))


def test_product_does_not_import_the_oracle():
    """oracle/ is test infrastructure: nothing under np-modeling_b200/ may import or execute it."""
    pkg = os.path.join(ROOT, 'np-modeling_b200')
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(base, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle', src, flags=re.M), os.path.join(base, f)
                assert 'np_oracle' not in src or f == '_none_', os.path.join(base, f)
