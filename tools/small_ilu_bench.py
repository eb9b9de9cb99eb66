#!/usr/bin/env python3
"""mat900 / mat10000 with ILU0 through the host entry point: us per iteration (BASELINE config 2, latency-bound)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
import numpy as np
cm = ge.load_package()
for nm in ("mat900", "mat10000"):
    m, _, ia, ja, a = cm.load_mm(os.path.join(ROOT, "tests", "golden", nm + ".mtx"))
    rs = np.random.RandomState(0)
    b = 1.0 + 4.0 * rs.rand(m)
    for mode, fn in (("ilu0", cm.bicgstab_lu_precond), ("unprec", cm.bicgstab)):
        best = None
        for _ in range(5):
            x, dt, st = fn(a, ia, ja, b, maxit=2000, tol=1e-6)
            best = dt if best is None else min(best, dt)
        print("%s %s: iterations=%d loop=%.3f ms  %.1f us/iteration launches=%d graph=%d" % (nm, mode, st["iterations"], best * 1e3, best * 1e6 / max(st["iterations"], 1), st["kernel_launches"], st["graph_replay"]))
