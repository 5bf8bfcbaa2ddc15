"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol the header declares.
No compute calls (there is no GPU in the build container, and the library has no CPU fallback)."""
import ctypes as C
import os

import pytest

from nori_b200 import abi


def test_library_exports_every_declared_symbol():
    syms = abi.declared_symbols()
    assert len(syms) >= 20
    L = C.CDLL(abi.LIB_PATH)
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing


def test_abi_version_and_struct_sizes():
    L = abi.lib()
    assert L.nb_abi_version() == 2      # 2: N GPUs behind the boundary (nb_create_multi, nb_comm_*, nb_render_gather)
    assert C.sizeof(abi.BsdfDesc) == 32
    assert C.sizeof(abi.EmitterDesc) == 16
    assert C.sizeof(abi.IntegratorDesc) == 16
    assert abi.RAY_DTYPE.itemsize == 32 and abi.HIT_DTYPE.itemsize == 20
    assert C.sizeof(abi.Stats) == 80


def test_no_cpu_fallback_without_gpu():
    """On a box without a CUDA device nb_create must FAIL (loudly), not fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    L = abi.lib()
    h = L.nb_create(0)
    assert not h
    assert b"no CUDA device" in L.nb_last_error()
    with pytest.raises(abi.NoriError):
        abi.Context(0)
    with pytest.raises(abi.NoriError, match="no CUDA device"):
        abi.Context([0, 1])            # the multi-GPU constructor fails the same way


def test_product_does_not_import_oracle():
    """The product package must never route through oracle/ (test infrastructure only)."""
    root = os.path.dirname(os.path.abspath(abi.__file__))
    for dp, _, fs in os.walk(root):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dp, f), errors="ignore").read()
                assert "pyoracle" not in src and "liboracle" not in src and "oracle.h" not in src, os.path.join(dp, f)
