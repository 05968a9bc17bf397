"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every symbol the header
declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from nmch_b200 import _build, capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    _build.build(only_if_missing=True)
    return capi.load()


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "nmch_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nmch_[a-z_]+)\s*\(", text)))


def test_header_symbols_all_exported(lib):
    names = _declared_symbols()
    assert len(names) >= 16
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/nmch_b200.h but not exported"
    assert set(capi.EXPORTS) == set(names)


def test_struct_layout_matches_header():
    # 2 int + 8 float + 5 int (+pad) + 3 u64 + 2 int
    assert C.sizeof(capi.NmchParams) == 96
    assert C.sizeof(capi.NmchMoments) == 32


def test_status_strings(lib):
    assert lib.nmch_status_string(0) == b"ok"
    assert b"CUDA" in lib.nmch_status_string(capi.ERR_CUDA)
    assert lib.nmch_version().startswith(b"nmch_b200")


def test_argument_validation_needs_no_gpu(lib):
    h = C.c_void_p()
    p = capi.NmchParams(512, 512, 1.0, 1.0, 0.1, 0.0, 0.5, -0.7, 0.1, 0.3, 0, 0, 0, 0, -1, 0, 0, 0, 0, 0)
    assert lib.nmch_engine_create(C.byref(p), C.byref(h)) == capi.ERR_ARG          # N = 0
    p.N = 1000
    p.method = 7
    assert lib.nmch_engine_create(C.byref(p), C.byref(h)) == capi.ERR_ARG
    p.method = 0
    p.first_path = 100                                                               # not a multiple of 4096
    assert lib.nmch_engine_create(C.byref(p), C.byref(h)) == capi.ERR_ARG
    assert b"4096" in lib.nmch_last_error()
    p.method = 1                                                                     # native EM: same alignment rule
    assert lib.nmch_engine_create(C.byref(p), C.byref(h)) == capi.ERR_ARG            # (block-uniform high counter word)
    assert b"4096" in lib.nmch_last_error()
    p.method, p.first_path, p.rng = 1, 0, 4                                          # the dense stream is an FE mode
    assert lib.nmch_engine_create(C.byref(p), C.byref(h)) == capi.ERR_ARG
    assert b"FE stream mode" in lib.nmch_last_error()


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    p = capi.NmchParams(512, 512, 1.0, 1.0, 0.1, 0.0, 0.5, -0.7, 0.1, 0.3, 1000, 0, 0, 0, -1, 0, 0, 0, 0, 0)
    assert lib.nmch_engine_create(C.byref(p), C.byref(h)) == capi.ERR_CUDA
    assert b"no CPU fallback" in lib.nmch_last_error()
    with pytest.raises(capi.NmchError):
        from nmch_b200 import Engine
        Engine()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "nmch_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("no oracle", ""), f"{f} mentions the oracle"
