"""Host-side C++ of the library that reads untrusted files (tpl_loader.cpp: text pair and binary container) and the f(T_k)
helpers (tpl_ftk.cpp), built with AddressSanitizer + UndefinedBehaviorSanitizer and driven over well-formed, truncated,
bit-flipped and malformed inputs (tests/host_harness/loader_harness.cpp).  No GPU, no CUDA: the error slot is stubbed."""
import glob
import os
import shutil
import subprocess

import pytest

import helpers

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "two_pass_lanczos_b200", "csrc")


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
def test_loader_and_ftk_under_asan_ubsan(tmp_path):
    exe = str(tmp_path / "harness")
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=all", "-fno-omit-frame-pointer",
           "-I", CSRC, "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "host_harness", "loader_harness.cpp"),
           os.path.join(CSRC, "tpl_loader.cpp"), os.path.join(CSRC, "tpl_ftk.cpp"), "-o", exe, "-pthread"]
    build = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    if build.returncode != 0 and "sanitize" in build.stderr and "cannot find" in build.stderr:
        pytest.skip("sanitizer runtimes not installed")
    assert build.returncode == 0, build.stderr[-3000:]
    dmx = sorted(glob.glob(os.path.join(helpers.GOLDEN, "netgen1000", "*.dmx")))[0]
    work = tmp_path / "work"
    work.mkdir()
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=1:abort_on_error=0", UBSAN_OPTIONS="print_stacktrace=1")
    env.pop("LD_PRELOAD", None)
    run = subprocess.run([exe, str(work), dmx, dmx[:-3] + "lines.qfc"], capture_output=True, text=True, timeout=600, env=env)
    assert run.returncode == 0, (run.stdout[-2000:], run.stderr[-4000:])
    assert "harness ok" in run.stdout
