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


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """The drop-in boundary is a C ABI: include/tplanczos.h compiles as strict C99 (no C++-isms, no torch / CUDA types) and
    a C program links against libtplanczos.so and resolves its entry points (host-only calls: no GPU needed)."""
    import shutil
    import subprocess

    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    src = tmp_path / "c_client.c"
    src.write_text('#include <stdio.h>\n#include <string.h>\n#include "tplanczos.h"\n'
                   "int main(void) {\n"
                   "  double al[2] = {2.0, 3.0}, be[1] = {1.0}, y[2], res[2];\n"
                   "  size_t ylen = 0;\n"
                   "  tpl_kkt* k = NULL;\n"
                   '  if (tpl_load_kkt("/nonexistent.dmx", "/nonexistent.qfc", &k) != TPL_ERR_IO) return 1;\n'
                   "  if (strlen(tpl_last_error_message()) == 0) return 2;\n"
                   "  if (tpl_ftk_inv(al, 2, be, 1, y, &ylen, NULL) != TPL_OK || ylen != 2) return 3;\n"
                   "  if (tpl_ftk_inv_residuals(al, 2, be, 1, 1.0, res) != TPL_OK) return 4;\n"
                   '  printf("%s %.17g %.17g\\n", tpl_version(), y[0], res[0]);\n'
                   "  return 0;\n}\n")
    exe = tmp_path / "c_client"
    libdir = os.path.dirname(_lib.LIB_PATH)
    build = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                            str(src), "-o", str(exe), "-L", libdir, "-ltplanczos", f"-Wl,-rpath,{libdir}"],
                           capture_output=True, text=True)
    assert build.returncode == 0, build.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True)
    assert run.returncode == 0, (run.returncode, run.stdout, run.stderr)
    y0, r0 = run.stdout.split()[-2:]
    assert float(y0) == pytest.approx(3.0 / 5.0, rel=1e-15)  # T = [[2,1],[1,3]], y = T^{-1} e1
    assert float(r0) == pytest.approx(0.5, rel=1e-15)        # ||b|| beta_1 / |alpha_1|
