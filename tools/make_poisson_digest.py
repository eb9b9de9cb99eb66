#!/usr/bin/env python3
"""One-off: run the CPU oracle on the Poisson N^3 system of BASELINE.json (x0 = ones, b = A x_true, tol 1e-10) and
commit its digest (iterations, sha256 of x and of the residual history) so that -m gpu tests and bench.py can assert
bit-identity at the size the metric is quoted on without running the oracle on the GPU box (256^3: ~10 min of CPU).

    python tools/make_poisson_digest.py 256 [ilu0]   ->  tests/golden/poisson<N>[_ilu0]_oracle_digest.json
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    ilu0 = len(sys.argv) > 2 and sys.argv[2] == "ilu0"
    O = ge.load_oracle()
    t0 = time.time()
    ia, ja, a = O.poisson3d(N)
    xt = O.xtrue(1234, 0, N ** 3)
    b = O.spmv(ia, ja, a, xt)
    if ilu0:
        x, st = O.bicgstab_ilu0(ia, ja, a, b, maxit=5000, tol=1e-10)
    else:
        x, st = O.bicgstab_unprec(ia, ja, a, b, maxit=5000, tol=1e-10)
    out = {"workload": "poisson3d_%d" % N, "mode": "ilu0" if ilu0 else "plain", "tol": 1e-10, "x0": "ones",
           "b": "A*x_true, x_true=hash(1234,i)", "iterations": int(st["iterations"]), "converged": bool(st["converged"]),
           "nrm_r0": float(st["nrm_r0"]), "nrm_r": float(st["nrm_r"]),
           "x_sha256": hashlib.sha256(np.ascontiguousarray(x).tobytes()).hexdigest(),
           "hist_sha256": hashlib.sha256(np.ascontiguousarray(st["hist"]).tobytes()).hexdigest(),
           "hist_len": int(len(st["hist"])), "b_sha256": hashlib.sha256(np.ascontiguousarray(b).tobytes()).hexdigest(),
           "rel_err_vs_xtrue": float(np.linalg.norm(x - xt) / np.linalg.norm(xt)),
           "oracle_seconds": time.time() - t0, "generated_by": "tools/make_poisson_digest.py (oracle/oracle.c)"}
    p = os.path.join(ROOT, "tests", "golden", "poisson%d%s_oracle_digest.json" % (N, "_ilu0" if ilu0 else ""))
    json.dump(out, open(p, "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
