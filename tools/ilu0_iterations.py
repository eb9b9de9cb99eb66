#!/usr/bin/env python3
"""A few ILU0-BiCGSTAB iterations on Poisson N^3 (default 256): the workload of the ncu launch list profiles/r2i_*."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
import torch
cm = ge.load_package()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
maxit = int(sys.argv[2]) if len(sys.argv) > 2 else 6
n, nnz = N ** 3, cm.poisson3d_nnz(N)
f64 = dict(dtype=torch.float64, device="cuda")
ia = torch.empty(n + 1, dtype=torch.int32, device="cuda"); ja = torch.empty(nnz, dtype=torch.int32, device="cuda"); a = torch.empty(nnz, **f64)
cm.gen_poisson3d_device(N, 0, n, ia.data_ptr(), ja.data_ptr(), a.data_ptr())
s = cm.Solver(n)
s.set_csr_device(nnz, a.data_ptr(), ia.data_ptr(), ja.data_ptr())
s.analyze(cm.MODE_ILU0)
xt = torch.empty(n, **f64); cm.gen_xtrue_device(1234, 0, n, xt.data_ptr())
b = torch.empty(n, **f64); s.spmv(xt.data_ptr(), b.data_ptr())
x = torch.zeros(n, **f64)
st = s.solve(cm.MODE_ILU0, b.data_ptr(), x.data_ptr(), maxit=maxit, tol=1e-30)
torch.cuda.synchronize()
print("iterations", st["iterations"], "loop ms", st["t_loop"] * 1e3, "sweep blocks", s.sweep_blocks())
