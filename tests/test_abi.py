"""The C-ABI shared library loads on a CPU-only box, exports every symbol include/tplanczos.h declares, and
FAILS LOUDLY (no CPU fallback) when asked to compute without a GPU."""
import os
import re

import numpy as np
import pytest

from two_pass_lanczos_b200 import _lib
from two_pass_lanczos_b200.error import CudaError
from two_pass_lanczos_b200.operators import LinOp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "tplanczos.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tpl_[a-z0-9_]+)\s*\(", text)) - {"tpl_step_callback", "tpl_ftk_solver"})


def test_every_declared_symbol_is_exported_and_bound():
    lib = _lib.load()
    names = header_symbols()
    assert len(names) >= 30
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/tplanczos.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert set(_lib.SIGNATURES) == set(names)


def test_no_torch_types_in_the_abi():
    text = open(os.path.join(ROOT, "include", "tplanczos.h")).read()
    assert "torch" not in text.lower() and "at::" not in text and "#include <cuda" not in text


def test_version_and_error_slot():
    lib = _lib.load()
    assert b"sm_100a" in lib.tpl_version()
    assert isinstance(lib.tpl_last_error_message(), bytes)


def _have_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        return False


@pytest.mark.skipif(_have_gpu(), reason="checks the behaviour on a box WITHOUT a GPU")
def test_compute_without_gpu_fails_loudly():
    with pytest.raises(CudaError) as e:
        LinOp.from_dense(np.eye(3))
    assert "no CPU fallback" in str(e.value)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under the product package may reference it."""
    pkg = os.path.join(ROOT, "two_pass_lanczos_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in text.lower(), f"{f} mentions the oracle"
