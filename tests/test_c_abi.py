"""The drop-in boundary from plain C: include/kr_engine.h must compile as C99 (cgo reads it with a C compiler) and a C
program must be able to drive a whole epoch through libkrengine.so without Python or C++ (tests/c_abi_smoke.c)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c_abi_smoke.c")


def _build(tmp_path):
    from kuberay_b200 import engine
    lib = engine.LIB_PATH
    assert os.path.exists(lib), "build the engine first (__graft_entry__.build())"
    exe = str(tmp_path / "c_abi_smoke")
    cmd = ["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), SRC, "-o", exe,
           lib, "-Wl,-rpath," + os.path.dirname(lib), "-Wl,--allow-shlib-undefined"]
    subprocess.check_call(cmd)
    return exe


def test_header_is_c99_and_the_library_links_from_c(tmp_path):
    """CPU box: the program compiles with -std=c99 -pedantic -Werror, links against the library and reports the documented
    refusal when no device is visible (no compute call happens without a GPU)."""
    exe = _build(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "no CUDA device" in out.stdout or "C ABI smoke: OK" in out.stdout
    assert "host builders from C: OK" in out.stdout          # kr_pod_build needs no device


@pytest.mark.gpu
def test_one_epoch_driven_from_plain_c(tmp_path):
    exe = _build(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    sys.stdout.write(out.stdout)
    assert out.returncode == 0 and "C ABI smoke: OK" in out.stdout, out.stdout + out.stderr
