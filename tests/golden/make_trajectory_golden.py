"""Long-trajectory and restart fixtures from the REFERENCE'S OWN unmodified program (oracle/_ref/topopt_ref = the reference
sources on oracle/petsc_lite, exact solves; see make_reference_golden.py).

1. tests/golden/trajectory_48x24x24.json -- BASELINE configs[0] (48x24x24 cantilever, volfrac 0.12, rmin 0.08, density
   filter) for 60 design iterations: the history main.cc:108-110 prints (6 decimals) and the final design (every 5th
   value of x at full precision, sums, state norm).  north_star: "final compliance within 1e-6 relative of the reference".
2. tests/golden/ref_restart_17x9x9/ -- the restart files the reference WRITES after 3 iterations of a 16x8x8 cantilever
   (Restart0x.dat: PETSc binary Vec stream x, xPhys?, xo1, xo2, U, L as TopOpt.cc:556-562 orders them; Restart0x_itr_f0.dat;
   RestartSol0x.dat), plus the history of the reference CONTINUING from them to iteration 6.  The product must load these
   files and print the same continuation (tests/test_dropin_gpu.py).

Usage (container with /root/reference):  make -C oracle && python tests/golden/make_trajectory_golden.py [--from DIR]
  --from DIR  ingest an existing run directory of case 1 (run.log + restart files) instead of re-running (minutes)."""
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_reference_golden import REF, ROOT, parse_history, read_petsc_vecs  # noqa: E402

TRAJ_ARGS = ["-nx", "49", "-ny", "25", "-nz", "25", "-nlvls", "4", "-maxItr", "60", "-rmin", "0.08", "-filter", "1", "-restart", "1"]


def final_restart(d):
    """the restart pair written last (the reference alternates 00 / 01, TopOpt.cc:525-534)"""
    cands = []
    for k in ("00", "01"):
        p = os.path.join(d, f"Restart{k}_itr_f0.dat")
        if os.path.exists(p):
            itr, fscale = open(p).read().split()
            cands.append((int(itr), float(fscale), k))
    return max(cands)


def trajectory(src=None):
    if src is None:
        src = tempfile.mkdtemp()
        with open(os.path.join(src, "run.log"), "w") as f:
            subprocess.run([REF] + TRAJ_ARGS, cwd=src, stdout=f, stderr=subprocess.STDOUT, check=True, timeout=4 * 3600)
    hist = parse_history(open(os.path.join(src, "run.log")).read())
    assert len(hist) == 60, len(hist)
    itr, fscale, k = final_restart(src)
    vecs = read_petsc_vecs(os.path.join(src, f"Restart{k}.dat"))
    usol = read_petsc_vecs(os.path.join(src, f"RestartSol{k}.dat"))[0]
    x = vecs[0]
    out = dict(generator="tests/golden/make_trajectory_golden.py", binary="oracle/_ref/topopt_ref (reference sources + oracle/petsc_lite, exact solves)",
               args=TRAJ_ARGS, shape=[49, 25, 25], history=hist, restart_itr=itr, restart_fscale=fscale, stride=5,
               x=[float(v) for v in x[::5]], x_sum=float(x.sum()), x_size=int(x.size), U_norm=float(np.linalg.norm(usol)))
    p = os.path.join(ROOT, "tests", "golden", "trajectory_48x24x24.json")
    json.dump(out, open(p, "w"), indent=0, separators=(",", ":"))
    print("wrote", p, os.path.getsize(p), "bytes; final fx", hist[-1][1])


def restart_fixture():
    base = ["-nx", "17", "-ny", "9", "-nz", "9", "-nlvls", "1", "-rmin", "0.3", "-restart", "1"]
    d = tempfile.mkdtemp()
    r = subprocess.run([REF] + base + ["-maxItr", "3"], cwd=d, capture_output=True, text=True, check=True)
    first = parse_history(r.stdout)
    itr, fscale, k = final_restart(d)
    assert itr == 3
    dst = os.path.join(ROOT, "tests", "golden", "ref_restart_17x9x9")
    os.makedirs(dst, exist_ok=True)
    for name in (f"Restart{k}.dat", f"Restart{k}_itr_f0.dat", f"RestartSol{k}.dat"):
        shutil.copy(os.path.join(d, name), os.path.join(dst, name.replace(k, "00")))
    # the reference continuing from its own files
    d2 = tempfile.mkdtemp()
    for name in os.listdir(dst):
        if name.endswith(".dat"):
            shutil.copy(os.path.join(dst, name), os.path.join(d2, "in_" + name))
    r2 = subprocess.run([REF] + base + ["-maxItr", "6", "-restartFileVec", "in_Restart00.dat", "-restartFileItr", "in_Restart00_itr_f0.dat",
                                        "-restartFileVecSol", "in_RestartSol00.dat"], cwd=d2, capture_output=True, text=True, check=True)
    cont = parse_history(r2.stdout)
    assert "Successful restart" in r2.stdout and [int(h[0]) for h in cont] == [4, 5, 6], r2.stdout[-1500:]
    json.dump(dict(generator="tests/golden/make_trajectory_golden.py", args=base, first_run_history=first, continued_history=cont,
                   note="files written by the reference after iteration 3; continued_history = the reference restarted from them"),
              open(os.path.join(dst, "history.json"), "w"), indent=1)
    print("wrote", dst, [h[1] for h in first], "->", [h[1] for h in cont])


if __name__ == "__main__":
    if not os.path.exists(REF):
        sys.exit("oracle/_ref/topopt_ref missing: run `make -C oracle` where /root/reference exists")
    src = sys.argv[sys.argv.index("--from") + 1] if "--from" in sys.argv else None
    restart_fixture()
    trajectory(src)
