"""The reference's two host programs, built from this repo's C++ host mirror (scripts/build_host.sh, no CMake)
and run on the GPU: `TestSimulation N K M` (host pointers through extern "C" MatrixMultiplicationKernel,
test/TestSimulation.cpp:66) and `RunHardware.exe N K M hw on` (Context / Buffer / Kernel, host/RunHardware.cpp).
Both verify against the host ReferenceImplementation with the reference's criterion and print its sentences.
(The file sorts last on purpose: it is the only GPU test that depends on a host compiler at run time.)"""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _build(tmp, *cfg):
    if not shutil.which("g++"):
        pytest.skip("no host compiler on this box")
    out = str(tmp)
    # the plain single-GPU programs (no NCCL driver linked in): the configuration of profiles/r01_host_executables.log
    r = subprocess.run(["bash", os.path.join(ROOT, "scripts", "build_host.sh"), out, *cfg], capture_output=True, text=True,
                       env=dict(os.environ, MM_HOST_NO_NCCL="1"))
    if r.returncode != 0:
        pytest.skip("host executables did not build here: " + (r.stdout + r.stderr)[-300:])
    return out


def _run(exe, *args):
    return subprocess.run([exe, *map(str, args)], capture_output=True, text=True, timeout=600)


def test_float_host_programs_verify_on_the_gpu(mm, tmp_path):
    out = _build(tmp_path)
    r = _run(os.path.join(out, "TestSimulation"), 513, 528, 528)        # the reference's CTest shape
    assert r.returncode == 0, r.stdout + r.stderr
    assert "successfully verified" in r.stdout
    r = _run(os.path.join(out, "RunHardware"), 1024, 1024, 1024, "hw", "on")
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Successfully verified." in r.stdout
    m = re.search(r"Kernel executed in ([0-9.e+-]+) seconds, corresponding to a performance of ([0-9.e+-]+) GOp/s", r.stdout)
    assert m, r.stdout                                                   # the line scripts/build_manager.py:601 parses
    assert float(m.group(2)) == pytest.approx(1e-9 * 2.0 * 1024 ** 3 / float(m.group(1)), rel=1e-3)
