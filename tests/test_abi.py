"""The C-ABI library builds without a GPU, loads, and exports every symbol include/eals_b200.h
declares; without a device every working entry point fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "eals_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(eals_[a-z_A-Z0-9]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    from eals_cpp_b200 import _lib
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == names
    assert lib.eals_abi_version() == 1


def test_params_struct_layout_and_defaults():
    from eals_cpp_b200 import _lib
    lib = _lib.load()
    p = _lib.EalsParams()
    lib.eals_default_params(C.byref(p))
    assert p.struct_bytes == C.sizeof(_lib.EalsParams)
    # run.sh defaults, main.cpp:133-144
    assert (p.factors, p.topk, p.w0, p.alpha, p.reg, p.init_mean, p.init_stdev) == (64, 10, 10.0, 0.75, 0.01, 0.0, 0.01)


def test_library_is_sm100a_only():
    import subprocess
    from eals_cpp_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("box has a GPU")
    from eals_cpp_b200._lib import EalsError
    from eals_cpp_b200.model import MF_fastALS, SparseMat
    sm = SparseMat.from_csr(2, 2, np.array([0, 1, 2], np.int64), np.array([0, 1], np.int32))
    with pytest.raises(EalsError, match="no CUDA device|CUDA"):
        MF_fastALS(sm, None, factors=8, device=0)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "eals_cpp_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dp, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "eals_oracle" not in src and "libeals_ref" not in src, f
