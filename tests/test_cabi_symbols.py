"""
The C-ABI shared library loads on a CPU-only machine and exports every symbol that
include/verde_b200.h declares (no compute calls here: there is no GPU).
"""
import ctypes
import os
import re

import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "verde_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vb200_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from verde_b200 import _cabi

    if not os.path.exists(_cabi.LIB_PATH):
        import __graft_entry__

        __graft_entry__.build()
    return _cabi.load()


def test_header_and_binding_agree(lib):
    from verde_b200 import _cabi

    assert declared_symbols() == sorted(_cabi.SIGNATURES)


def test_every_declared_symbol_is_exported(lib):
    for name in declared_symbols():
        assert hasattr(lib, name), name


def test_fit_info_layout_matches_header():
    from verde_b200 import _cabi

    text = open(HEADER).read()
    body = re.search(r"typedef struct vb200_fit_info \{(.*?)\} vb200_fit_info;", text, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"(int32_t|int64_t|double)\s+([a-z_]+)\s*;", body)
    ctype = {"int32_t": ctypes.c_int32, "int64_t": ctypes.c_int64, "double": ctypes.c_double}
    assert [(n, ctype[t]) for t, n in fields] == list(_cabi.FitInfo._fields_)


def test_library_answers_without_a_gpu(lib):
    assert lib.vb200_version() >= 100
    assert lib.vb200_gram_doubles(128) == 128 * 128
    assert lib.vb200_gram_doubles(129) == 3 * 128 * 128
    assert lib.vb200_stats_doubles(200) == 4 * 256
    assert lib.vb200_set_option(b"no_such_option", 1.0) == -1
    assert b"unknown option" in lib.vb200_last_error()


def test_product_path_fails_loudly_without_cuda(lib):
    """no silent CPU fallback: on a machine without a GPU the estimators raise"""
    import numpy as np

    import verde_b200 as vb

    if lib.vb200_device_count() > 0:
        pytest.skip("a CUDA device is present")
    e = np.linspace(0, 1, 10)
    with pytest.raises(vb.EngineError, match="no CPU fallback"):
        vb.Spline(damping=1e-3).fit((e, e), e)
    with pytest.raises(vb.EngineError):
        vb.Spline(damping=1e-3).jacobian((e, e), (e, e))


def test_product_does_not_import_the_oracle():
    """only tests/, smoke() and bench.py's CPU legs may touch oracle/"""
    pkg = os.path.join(ROOT, "verde_b200")
    for dirpath, _, files in os.walk(pkg):
        for name in files:
            if name.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, name)).read()
                assert "verde_oracle" not in text and "greens_oracle" not in text, name
