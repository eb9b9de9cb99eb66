#!/usr/bin/env python3
"""CPU oracle's block-Jacobi ILU(0)-BiCGSTAB on Poisson N^3 for the row partition of `world` ranks (cudamat_partition_rows):
iteration count, relative error and sha256 of x — what the sharded GPU solve of tests/dist_gpu_worker.py must give.
usage: tools/oracle_block_ilu0.py N world [world ...]   (256^3 takes ~10 minutes per partition on one core)"""
import hashlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import __graft_entry__ as ge  # noqa: E402
O = ge.load_oracle(); cm = ge.load_package()
N = int(sys.argv[1]); n = N ** 3
ia, ja, a = O.poisson3d(N)
xt = O.xtrue(1234, 0, n); b = O.spmv(ia, ja, a, xt)
for world in [int(v) for v in sys.argv[2:]]:
    rs = [0] + [cm.partition_rows(n, world, r)[1] for r in range(world)]
    t = time.time()
    x, st = O.bicgstab_ilu0_blocks(ia, ja, a, b, rs, maxit=5000, tol=1e-10)
    print(json.dumps({"N": N, "world": world, "row_starts": rs, "iterations": st["iterations"], "converged": bool(st["converged"]),
                      "rel_err_vs_xtrue": float(np.linalg.norm(x - xt) / np.linalg.norm(xt)),
                      "x_sha256": hashlib.sha256(np.ascontiguousarray(x).tobytes()).hexdigest(), "seconds": round(time.time() - t, 1)}), flush=True)
