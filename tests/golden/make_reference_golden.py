"""Generate tests/golden/reference_runs.json from the REFERENCE'S OWN unmodified program.

oracle/_ref/topopt_ref is main.cc + TopOpt.cc + LinearElasticity.cc + Filter.cc + PDEFilter.cc + MMA.cc + MPIIO.cc of the
reference, compiled where they lie under /root/reference (oracle/Makefile) against oracle/petsc_lite -- a serial
restatement of the PETSc/MPI slice they use, with KSPSolve replaced by an exact Jacobi-PCG solve.  This script runs it
on a few small cantilevers covering the three filter types and the Heaviside projection and stores
  * the optimisation history exactly as main.cc:108-110 prints it,
  * the final design x / xPhys (first two Vecs of the reference's restart file, TopOpt.cc:556-562) at full precision
    (every value for grids up to 600 elements, every 3rd beyond), the state vector's norm and the restart scalars.
The fixtures travel to the GPU box (where /root/reference does not exist); tests/test_reference_cpu.py checks the numpy
oracle and -- where the binary exists -- the binary itself against them.

With --big it also runs the two grids of the surveyor's known answers (48x24x24 and the reference default
64x32x32, uniform density 0.12) for ONE design iteration and stores the compliance the reference prints (minutes of CPU
time and gigabytes of std::map matrix: petsc_lite is built for fidelity, not speed); without it the stored values are kept.

Usage (in the container that has /root/reference):  make -C oracle && python tests/golden/make_reference_golden.py [--big]
"""
import json
import os
import re
import struct
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.path.join(ROOT, "oracle", "_ref", "topopt_ref")
OUT = os.path.join(ROOT, "tests", "golden", "reference_runs.json")

CASES = [
    dict(name="density_17x9x9", shape=[17, 9, 9], iters=4, args=["-rmin", "0.3"], oracle=dict(rmin=0.3)),
    dict(name="sensitivity_17x9x9", shape=[17, 9, 9], iters=4, args=["-rmin", "0.3", "-filter", "0"],
         oracle=dict(rmin=0.3, filter_type=0)),
    dict(name="pde_13x9x7", shape=[13, 9, 7], iters=3, args=["-rmin", "0.25", "-filter", "2"], oracle=dict(rmin=0.25, filter_type=2)),
    dict(name="projection_13x9x7", shape=[13, 9, 7], iters=3,
         args=["-rmin", "0.3", "-projectionFilter", "1", "-beta", "2.0", "-eta", "0.5"],
         oracle=dict(rmin=0.3, projection=True, beta=2.0, eta=0.5)),
]
HIST_RE = re.compile(r"It\.: (\d+), True fx: ([-\d.e+]+), Scaled fx: ([-\d.e+]+), gx\[0\]: ([-\d.e+]+), ch\.: ([-\d.e+]+), mnd\.: ([-\d.e+]+)")


def parse_history(out):
    return [[float(v) for v in m.groups()] for m in (HIST_RE.match(line) for line in out.splitlines()) if m]


def read_petsc_vecs(path):
    """PETSc binary Vec stream: big-endian int32 classid 1211214, int32 n, n float64."""
    data, pos, out = open(path, "rb").read(), 0, []
    while pos < len(data):
        cid, n = struct.unpack(">ii", data[pos:pos + 8])
        assert cid == 1211214
        out.append(np.frombuffer(data[pos + 8:pos + 8 + 8 * n], dtype=">f8").astype(float))
        pos += 8 + 8 * n
    return out


def run_reference(case, workdir=None):
    d = workdir or tempfile.mkdtemp()
    nx, ny, nz = case["shape"]
    cmd = [REF, "-nx", str(nx), "-ny", str(ny), "-nz", str(nz), "-nlvls", "1", "-maxItr", str(case["iters"]), "-restart", "1"] + case["args"]
    r = subprocess.run(cmd, cwd=d, capture_output=True, text=True, timeout=3600)
    assert r.returncode == 0, r.stdout[-800:] + r.stderr[-800:]
    return r.stdout, d


def sample_stride(n):
    return 1 if n <= 600 else 3


def main():
    if not os.path.exists(REF):
        sys.exit("oracle/_ref/topopt_ref missing: run `make -C oracle` where /root/reference exists")
    out = {"generator": "tests/golden/make_reference_golden.py", "binary": "oracle/_ref/topopt_ref (reference sources + oracle/petsc_lite)",
           "cases": []}
    for case in CASES:
        stdout, d = run_reference(case)
        hist = parse_history(stdout)
        assert len(hist) == case["iters"]
        vecs = read_petsc_vecs(os.path.join(d, "Restart00.dat"))  # after < 10 iterations only the final dump (flip -> file 00) exists
        usol = read_petsc_vecs(os.path.join(d, "RestartSol00.dat"))[0]
        itr, fscale = open(os.path.join(d, "Restart00_itr_f0.dat")).read().split()
        x, xphys = vecs[0], vecs[1]
        m = re.search(r"results in a stencil of (\d+) elements", stdout)  # Filter.cc:334 (cone filters only)
        st = sample_stride(x.size)
        out["cases"].append(dict(name=case["name"], shape=case["shape"], iters=case["iters"], args=case["args"], oracle=case["oracle"],
                                 history=hist, stride=st, x=[float(v) for v in x[::st]], xPhys=[float(v) for v in xphys[::st]],
                                 x_sum=float(x.sum()), xPhys_sum=float(xphys.sum()), U_norm=float(np.linalg.norm(usol)),
                                 U_size=int(usol.size), restart_itr=int(itr), restart_fscale=float(fscale), elemconn=int(m.group(1)) if m else None,
                                 container_bytes=os.path.getsize(os.path.join(d, "output_00000.dat"))))
        print(case["name"], "fx:", [h[1] for h in hist])
    known = {}
    if os.path.exists(OUT):
        known = json.load(open(OUT)).get("known_answers", {})
    if "--big" in sys.argv:
        for shape in ([49, 25, 25], [65, 33, 33]):
            stdout, _ = run_reference(dict(shape=shape, iters=1, args=[]))
            h = parse_history(stdout)
            known["x".join(str(v - 1) for v in shape)] = dict(iteration1_true_fx_printed=h[0][1], line=[l for l in stdout.splitlines() if l.startswith("It.: 1")][0])
            print(shape, h[0])
    out["known_answers"] = known
    with open(OUT, "w") as f:
        json.dump(out, f, indent=0, separators=(",", ":"))
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
