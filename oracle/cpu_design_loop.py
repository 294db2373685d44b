"""The reference's design iteration (main.cc:48-123) on the host cores, with the reference's OWN solver configuration
(FGMRES + Galerkin PCMG + GMRES/SOR smoothers, oracle/c/cpu_ref.c) -- TEST / BASELINE INFRASTRUCTURE ONLY: this is what
bench.py's `--impl reference` arm and its `cpu_baseline` leg time, and what tests/test_cpu_design_loop.py checks against
the numpy oracle.  Every step is a WHOLE design iteration: CSR re-assembly (LinearElasticity.cc:190), the Galerkin
re-setup (:198-200), the Krylov solve from the previous state (:204), objective/sensitivities (:403-437), the
sensitivity filter (Filter.cc:178-192), the MMA update (the reference's own MMA.cc as oracle/_ref/libmma_ref.so when it
was built, else the numpy restatement), the design filter (Filter.cc:66-71)."""
import time

import numpy as np

import cpu_ref
import mma_ref
import topopt_oracle as o


class CpuDesignLoop:
    def __init__(self, nex, ney, nez, volfrac=0.12, penal=3.0, Emin=1e-9, Emax=1.0, rmin=0.08, nlvls=4, rtol=1e-5,
                 Xmin=0.0, Xmax=1.0, movlim=0.2, nthreads=0, xc=(0, 2, 0, 1, 0, 1), nu=0.3, use_mma_ref=True):
        self.cr = cpu_ref.CpuRef(nex + 1, ney + 1, nez + 1, xc=xc, nu=nu, nthreads=nthreads)
        self.p = dict(volfrac=volfrac, penal=penal, Emin=Emin, Emax=Emax, rmin=rmin, nlvls=nlvls, rtol=rtol, Xmin=Xmin,
                      Xmax=Xmax, movlim=movlim)
        n = self.cr.nel
        self.x = np.full(n, volfrac)
        self.xold = self.x.copy()
        self.U = np.zeros(self.cr.ndof)
        self.mma_kind = "reference MMA.cc (oracle/_ref/libmma_ref.so, one process)" if (use_mma_ref and mma_ref.available()) \
            else "numpy restatement of MMA.cc (one process)"
        self.mma = mma_ref.MMARef(n, 1) if (use_mma_ref and mma_ref.available()) else o.MMAOracle(n, 1)
        self.xPhys, _ = self.cr.cone_filter(rmin, self.x, 0)  # main.cc:48 (filter 1, no projection: xPhys = xTilde)
        self.fscale = None
        self.itr = 0
        self.hist = []

    def step(self):
        p, cr = self.p, self.cr
        t0 = time.perf_counter()
        self.itr += 1
        t_asm = cr.assemble(self.xPhys, p["Emin"], p["Emax"], p["penal"])
        t_setup = cr.mg_setup(p["nlvls"])
        self.U, its, rel, t_solve = cr.fgmres_mg(self.U, rtol=p["rtol"], maxits=200, restart=100)
        fx, gx, dfdx, t_obj = cr.objective(self.U, self.xPhys, p["Emin"], p["Emax"], p["penal"], p["volfrac"])
        if self.itr == 1:
            self.fscale = 10.0 / fx  # main.cc:68-70
        dfdx *= self.fscale
        t1 = time.perf_counter()
        dfdx, _ = cr.cone_filter(p["rmin"], dfdx, 1)  # Filter.cc:178-192
        dgdx, _ = cr.cone_filter(p["rmin"], np.full(cr.nel, 1.0 / cr.nel), 1)
        t2 = time.perf_counter()
        if isinstance(self.mma, o.MMAOracle):
            xmin, xmax = self.mma.set_outer_movelimit(p["Xmin"], p["Xmax"], p["movlim"], self.x)
            xnew = self.mma.update(self.x, dfdx, gx, [dgdx], xmin, xmax)
            ch = float(np.abs(xnew - self.xold).max())
            self.xold = xnew.copy()
        else:
            xmin, xmax = self.mma.set_outer_movelimit(p["Xmin"], p["Xmax"], p["movlim"], self.x)
            xnew = self.mma.update(self.x, dfdx, gx, [dgdx], xmin, xmax)
            ch = self.mma.design_change(xnew, self.xold)
        self.x = xnew
        t3 = time.perf_counter()
        self.xPhys, _ = cr.cone_filter(p["rmin"], self.x, 0)  # main.cc:98
        mnd = float(o.mnd(self.xPhys))
        t4 = time.perf_counter()
        rec = dict(itr=self.itr, fx=fx, gx=gx, ch=ch, mnd=mnd, krylov_its=its, rel_residual=rel,
                   seconds=dict(total=t4 - t0, assembly=t_asm, galerkin_setup=t_setup, solve=t_solve, objective=t_obj,
                                filters=(t2 - t1) + (t4 - t3), mma=t3 - t2))
        self.hist.append(rec)
        return rec

    def close(self):
        if hasattr(self.mma, "close"):
            self.mma.close()
        self.cr.close()
