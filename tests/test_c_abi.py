"""The C ABI consumed from plain C (tests/c/abi_smoke.c): the header compiles as C11 with -Wall -Wextra -Werror, the program
links against libcova_b200.so alone (no Python, no CUDA headers) and runs.  Without a device it checks the no-fallback
contract and the host-only entry points; on the GPU box (marker gpu) it also runs the element shims on the device."""
import os
import subprocess

import pytest

from cova_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    _lib.load()                                                      # raises if the library has not been built
    exe = str(tmp_path / "abi_smoke")
    lib_dir = os.path.join(ROOT, "cova_b200")
    r = subprocess.run(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "tests", "c", "abi_smoke.c"), "-L" + lib_dir, "-lcova_b200", "-Wl,-rpath," + lib_dir, "-o", exe],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_c_program_links_and_runs(tmp_path):
    r = subprocess.run([_build(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.stdout, r.stderr)
    assert "abi_smoke ok" in r.stdout


@pytest.mark.gpu
def test_c_program_runs_the_elements_on_the_device(tmp_path):
    r = subprocess.run([_build(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.stdout, r.stderr)
    assert r.stdout.strip().endswith("abi_smoke ok")                 # the device branch, not the no-device one
