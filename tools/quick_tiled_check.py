#!/usr/bin/env python3
"""Seconds-long GPU sanity check without torch: smoke() (mat900, ILU0 + unpreconditioned, vs the oracle) and one
Poisson 48^3 solve through the host entry point (AUTO = TILED: full tiles, one ragged tile) that must reproduce the
oracle bit for bit."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

t0 = time.time()
ge.smoke()
cm, O = ge.load_package(), ge.load_oracle()
ia, ja, a = O.poisson3d(48)
n = len(ia) - 1
b = O.spmv(ia, ja, a, O.xtrue(1234, 0, n))
x, dt, st = cm.bicgstab(a, ia, ja, b, maxit=500, tol=1e-10)
xo, so = O.bicgstab_unprec(ia, ja, a, b, maxit=500, tol=1e-10)
print("poisson48: variant=%d gpu iters=%d oracle iters=%d bit-identical=%s  (%.1f s total)"
      % (st["spmv_variant"], st["iterations"], so["iterations"], np.array_equal(x, xo), time.time() - t0))
assert st["spmv_variant"] == cm.SPMV_TILED and st["iterations"] == so["iterations"] and np.array_equal(x, xo)
print("quick check OK")
