#!/usr/bin/env python3
"""Wall time of the one-shot host entry point on the reference's own fixtures (mat900 / mat10000, ILU0 and plain):
set-up (create, upload, analysis, ILU0, plans) against the iteration loop.  CUDAMAT_TIMING=1 adds the library's breakdown."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import __graft_entry__ as ge  # noqa: E402
cm = ge.load_package()
for nm in ("mat900", "mat10000"):
    m, _, ia, ja, a = cm.load_mm(os.path.join(ROOT, "tests", "golden", nm + ".mtx"))
    b = np.ones(m)
    for mode, fn in (("ilu0", cm.bicgstab_lu_precond), ("plain", cm.bicgstab)):
        for rep in range(4):
            t0 = time.perf_counter()
            x, dt, st = fn(a, ia, ja, b, maxit=2000, tol=1e-6)
            w = time.perf_counter() - t0
            print(f"{nm} {mode} call {rep}: wall {w*1e3:8.3f} ms, loop {dt*1e3:8.3f} ms, iterations {st['iterations']}, "
                  f"analysis {st['t_analysis']*1e3:.3f} ms, ilu0 {st['t_ilu0']*1e3:.3f} ms, h2d {st['t_h2d']*1e3:.3f} ms", flush=True)
