"""Host-array entry points with PAGEABLE (malloc'ed / numpy) arrays large enough for the threaded pinned staging of
csrc/hostcopy.cpp (>= 8 MB per array): same bits as the device-resident path and as pinned host arrays.
Reference callers pass malloc'ed arrays: example.cpp:96-104,252."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_pageable_host_arrays_take_the_staged_copy_and_give_the_same_bits(cm, torch_cuda):
    torch = torch_cuda
    N = 128
    n, nnz = N ** 3, cm.poisson3d_nnz(N)
    ia = torch.empty(n + 1, dtype=torch.int32, device="cuda")
    ja = torch.empty(nnz, dtype=torch.int32, device="cuda")
    a = torch.empty(nnz, dtype=torch.float64, device="cuda")
    cm.gen_poisson3d_device(N, 0, n, ia.data_ptr(), ja.data_ptr(), a.data_ptr())
    xt = torch.empty(n, dtype=torch.float64, device="cuda")
    cm.gen_xtrue_device(1234, 0, n, xt.data_ptr())
    s = cm.Solver(n)
    s.set_csr_device(nnz, a.data_ptr(), ia.data_ptr(), ja.data_ptr(), keep=(ia, ja, a))
    s.analyze(0)
    b = torch.empty(n, dtype=torch.float64, device="cuda")
    s.spmv(xt.data_ptr(), b.data_ptr())
    x = torch.zeros(n, dtype=torch.float64, device="cuda")
    st = s.solve(0, b.data_ptr(), x.data_ptr(), maxit=40, tol=1e-30)
    torch.cuda.synchronize()
    s.close()
    # pageable copies: (n+1)*4 = 8 MB + 4 bytes (a ragged last chunk), ja 58 MB, a 117 MB, b 16.8 MB, x back 16.8 MB
    h_ia, h_ja, h_a, h_b = (t.cpu().numpy().copy() for t in (ia, ja, a, b))
    assert h_ia.nbytes >= 8 << 20 and h_ia.nbytes % (4 << 20) == 4
    xh, dt, sth = cm.bicgstab(h_a, h_ia, h_ja, h_b, maxit=40, tol=1e-30)
    assert sth["iterations"] == st["iterations"] == 40
    assert np.array_equal(xh, x.cpu().numpy())
    # pinned host arrays (plain cudaMemcpyAsync) agree too
    pins = [torch.from_numpy(v).pin_memory() for v in (h_a, h_ia, h_ja, h_b)]
    xp, _, stp = cm.bicgstab(*(p.numpy() for p in pins), maxit=40, tol=1e-30)
    assert np.array_equal(xp, xh) and stp["nrm_r"] == sth["nrm_r"]
    # base-1 arrays through the same path (k_sub_base runs after the staged upload)
    x1, _, st1 = cm.bicgstab(h_a, h_ia + 1, h_ja + 1, h_b, maxit=40, tol=1e-30)
    assert np.array_equal(x1, xh)
    # staged download: the ILU0 factor (117 MB) through cudamat_ilu0_host equals the device copy of the factor
    M, lv, zp = cm.ilu0_host(h_a, h_ia, h_ja)
    s = cm.Solver(n)
    s.set_csr_device(nnz, a.data_ptr(), ia.data_ptr(), ja.data_ptr(), keep=(ia, ja, a))
    s.analyze(2)
    import ctypes as C
    Mp = torch.empty(nnz, dtype=torch.float64).pin_memory()               # pinned: plain cudaMemcpyAsync
    assert cm.lib.cudamat_get_ilu0_host(s.h, C.cast(Mp.data_ptr(), C.POINTER(C.c_double))) == 0
    Md = Mp.numpy()
    assert np.array_equal(s.ilu0_values(nnz), Md)                          # pageable numpy: staged download
    s.close()
    assert zp == 0 and lv == (3 * (N - 1) + 1, 3 * (N - 1) + 1)
    assert np.array_equal(M, Md)
