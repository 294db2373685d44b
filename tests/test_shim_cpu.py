"""CPU tests of the host shim (stub petsc.h / mpi.h): single process and the 2-process shared-memory mini-MPI
(the multi-rank host path; the device side talks NCCL and is covered by the -m gpu tests)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "topopt_in_petsc_b200", "host")
ARGS = ["-n", "1000", "-neg", "-2.5", "-off", "0", "-name", "abc"]


def build(tmp_path, with_mma):
    ref = "/root/reference"
    exe = str(tmp_path / ("selftest_mma" if with_mma else "selftest"))
    cmd = ["g++", "-O1", "-std=c++17", "-w", "-DTOB_SHIM_HOST_ONLY", f"-I{HOST}/include", "-o", exe,
           os.path.join(ROOT, "tests", "shim", "shim_selftest.cc"), os.path.join(HOST, "petsc_shim.cc")]
    if with_mma:
        cmd += ["-DWITH_REF_MMA", f"-I{ref}", os.path.join(ref, "MMA.cc")]
    subprocess.run(cmd, check=True)
    return exe


def run(exe, nranks, cwd):
    cmd = [exe] + ARGS if nranks == 1 else [os.path.join(HOST, "tobrun"), "-np", str(nranks), exe] + ARGS
    r = subprocess.run(cmd, cwd=cwd, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert f"SHIM_SELFTEST_OK ranks={nranks}" in r.stdout
    return r.stdout


@pytest.mark.parametrize("nranks", [1, 2, 3])
def test_shim_selftest(tmp_path, nranks):
    run(build(tmp_path, False), nranks, tmp_path)


def test_reference_mma_rank_invariance(tmp_path):
    """The reference's unmodified MMA.cc on the shim: 1-rank and 2-rank runs give the same design (its only
    communication is MPI_Allreduce, MMA.cc:420,626,764,834,925)."""
    if not os.path.exists("/root/reference/MMA.cc"):
        pytest.skip("reference checkout not present")
    exe = build(tmp_path, True)
    c1 = float(re.search(r"MMA_CHECKSUM (\S+)", run(exe, 1, tmp_path)).group(1))
    c2 = float(re.search(r"MMA_CHECKSUM (\S+)", run(exe, 2, tmp_path)).group(1))
    assert abs(c1 - c2) <= 1e-9 * abs(c1)
