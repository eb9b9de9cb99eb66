#!/usr/bin/env python3
"""Small end-to-end exercise of every kernel family for compute-sanitizer (memcheck): all SpMV variants, both solver modes,
ILU0 analysis / factorisation / sweeps (shared-memory and sync-free), generators, on small systems."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
import numpy as np
import torch
cm = ge.load_package()
O = ge.load_oracle()
ge.smoke()
for N in (20, 33):
    n = N ** 3
    nnz = cm.poisson3d_nnz(N)
    ia = torch.empty(n + 1, dtype=torch.int32, device="cuda"); ja = torch.empty(nnz, dtype=torch.int32, device="cuda")
    a = torch.empty(nnz, dtype=torch.float64, device="cuda")
    cm.gen_poisson3d_device(N, 0, n, ia.data_ptr(), ja.data_ptr(), a.data_ptr())
    xt = torch.empty(n, dtype=torch.float64, device="cuda"); cm.gen_xtrue_device(1, 0, n, xt.data_ptr())
    ref = None
    for v in (1, 2, 3, 4, 5):
        s = cm.Solver(n); s.set_option("spmv_variant", v)
        s.set_csr_device(nnz, a.data_ptr(), ia.data_ptr(), ja.data_ptr())
        s.analyze(0)
        b = torch.empty(n, dtype=torch.float64, device="cuda"); s.spmv(xt.data_ptr(), b.data_ptr(), variant=v)
        x = torch.zeros(n, dtype=torch.float64, device="cuda")
        st = s.solve(0, b.data_ptr(), x.data_ptr(), maxit=300, tol=1e-10)
        torch.cuda.synchronize()
        if ref is None: ref = (b.clone(), x.clone(), st["iterations"])
        assert torch.equal(b, ref[0]) and torch.equal(x, ref[1]) and st["iterations"] == ref[2], (N, v)
        s.close()
    for no_smem in (0, 1):
        s = cm.Solver(n); s.set_option("sptrsv_no_smem", no_smem)
        s.set_csr_device(nnz, a.data_ptr(), ia.data_ptr(), ja.data_ptr())
        s.analyze(2)
        x = torch.zeros(n, dtype=torch.float64, device="cuda")
        st = s.solve(2, ref[0].data_ptr(), x.data_ptr(), maxit=300, tol=1e-10)
        assert st["converged"], (N, no_smem)
        s.close()
# round 2 kernels: MARCH (plane-marching ring, all fuse modes), STREAM (one pass and column-blocked), the persistent cooperative
# iteration kernel (grid barrier and cluster barrier forms)
N = 64
n = N ** 3
nnz = cm.poisson3d_nnz(N)
ia = torch.empty(n + 1, dtype=torch.int32, device="cuda"); ja = torch.empty(nnz, dtype=torch.int32, device="cuda")
a = torch.empty(nnz, dtype=torch.float64, device="cuda")
cm.gen_poisson3d_device(N, 0, n, ia.data_ptr(), ja.data_ptr(), a.data_ptr())
xt = torch.empty(n, dtype=torch.float64, device="cuda"); cm.gen_xtrue_device(1, 0, n, xt.data_ptr())
ref = None
for fuse, persist in ((0, 0), (1, 0), (2, 0), (3, 0), (0, 1)):
    s = cm.Solver(n, stream=torch.cuda.current_stream().cuda_stream)
    s.set_option("fuse", fuse); s.set_option("persist", persist)
    s.set_csr_device(nnz, a.data_ptr(), ia.data_ptr(), ja.data_ptr())
    assert s.analyze(0)["spmv_variant"] == 6
    b = torch.empty(n, dtype=torch.float64, device="cuda"); s.spmv(xt.data_ptr(), b.data_ptr())
    x = torch.zeros(n, dtype=torch.float64, device="cuda")
    st = s.solve(0, b.data_ptr(), x.data_ptr(), maxit=12, tol=1e-10)
    torch.cuda.synchronize()
    if ref is None: ref = x.clone()
    assert torch.equal(x, ref) and st["iterations"] == 12, (fuse, persist)
    s.close()
m, _, mia, mja, ma = cm.load_mm(os.path.join(ROOT, "tests", "golden", "mat10000.mtx"))
xx, dt, st = cm.bicgstab(ma, mia, mja, np.ones(m), maxit=40, tol=1e-10)          # persistent kernel, one thread-block cluster
assert st["fused"] == 4 and st["iterations"] == 40
nr = 3000
ia = torch.empty(nr + 1, dtype=torch.int32, device="cuda")
nz = cm.gen_random_dd_device(nr, 5, ia.data_ptr())
ja = torch.empty(nz, dtype=torch.int32, device="cuda"); a = torch.empty(nz, dtype=torch.float64, device="cuda")
cm.gen_random_dd_device(nr, 5, ia.data_ptr(), ja.data_ptr(), a.data_ptr())
s = cm.Solver(nr); s.set_csr_device(nz, a.data_ptr(), ia.data_ptr(), ja.data_ptr()); s.analyze(2)
b = torch.ones(nr, dtype=torch.float64, device="cuda"); x = torch.zeros(nr, dtype=torch.float64, device="cuda")
st = s.solve(2, b.data_ptr(), x.data_ptr(), maxit=200, tol=1e-10); assert st["converged"]
st = s.solve(0, b.data_ptr(), x.data_ptr(), maxit=200, tol=1e-10); assert st["converged"]
xs = x.clone()
s.close()
for K in (1, 3):                                                                    # STREAM: one pass / 3 column blocks
    s = cm.Solver(nr); s.set_option("persist", 0); s.set_option("stream_blocks", K)
    s.set_csr_device(nz, a.data_ptr(), ia.data_ptr(), ja.data_ptr())
    assert s.analyze(0)["spmv_variant"] == 7
    st = s.solve(0, b.data_ptr(), x.data_ptr(), maxit=200, tol=1e-10)
    torch.cuda.synchronize()
    assert st["converged"] and torch.equal(x, xs), K
    s.close()
torch.cuda.synchronize()
print("sanitize_small OK")
