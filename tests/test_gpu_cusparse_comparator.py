"""Test-only comparator built on MODERN cuSPARSE: the reference's arithmetic lives in legacy
cuSPARSE calls that no longer exist (cusparseDcsrilu0, cusparseDcsrsv_*), their successors do.  cuSPARSE is loaded with
ctypes here and ONLY here — the product never links it.

  * ILU(0): cusparseDcsrilu02 (successor of cusparseDcsrilu0, pbicgstab.cu:359) must agree with cudamat_ilu0 on the same
    pattern to 1e-12 ("ILU0 factor bit-exact in structure, values to 1e-12");
  * L / U sweeps: cusparseSpSV (successor of cusparseDcsrsv_solve, pbicgstab.cu:94,98) vs cudamat_sptrsv_device."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _load_cusparse():
    import glob
    import torch
    cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cusparse", "lib", "libcusparse.so.12"))
    cands += ["/usr/local/cuda/lib64/libcusparse.so.12", "libcusparse.so.12"]
    for p in cands:
        try:
            return C.CDLL(p)
        except OSError:
            continue
    return None


def test_ilu0_and_sweeps_agree_with_modern_cusparse(cm, O, pin, torch_cuda):
    torch = torch_cuda
    cs = _load_cusparse()
    if cs is None:
        pytest.skip("libcusparse.so.12 not found")
    vp = C.c_void_p
    h = vp()
    assert cs.cusparseCreate(C.byref(h)) == 0
    mats = {"mat900": (pin["mat900_ia"] - 1, pin["mat900_ja"] - 1, pin["mat900_a"]),
            "mat10000": (pin["mat10000_ia"] - 1, pin["mat10000_ja"] - 1, pin["mat10000_a"]),
            "poisson24": O.poisson3d(24), "random_dd": O.random_dd(3000, 20240)}
    for nm, (ia, ja, a) in mats.items():
        n, nnz = len(ia) - 1, len(a)
        d_ia = torch.from_numpy(np.ascontiguousarray(ia, dtype=np.int32)).cuda()
        d_ja = torch.from_numpy(np.ascontiguousarray(ja, dtype=np.int32)).cuda()
        d_m = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()          # factored in place by cuSPARSE
        descr, info = vp(), vp()
        assert cs.cusparseCreateMatDescr(C.byref(descr)) == 0
        assert cs.cusparseCreateCsrilu02Info(C.byref(info)) == 0
        bs = C.c_int(0)
        assert cs.cusparseDcsrilu02_bufferSize(h, n, nnz, descr, vp(d_m.data_ptr()), vp(d_ia.data_ptr()), vp(d_ja.data_ptr()), info, C.byref(bs)) == 0
        buf = torch.empty(max(bs.value, 16), dtype=torch.uint8, device="cuda")
        POLICY_NO_LEVEL = 0
        assert cs.cusparseDcsrilu02_analysis(h, n, nnz, descr, vp(d_m.data_ptr()), vp(d_ia.data_ptr()), vp(d_ja.data_ptr()), info, POLICY_NO_LEVEL, vp(buf.data_ptr())) == 0
        assert cs.cusparseDcsrilu02(h, n, nnz, descr, vp(d_m.data_ptr()), vp(d_ia.data_ptr()), vp(d_ja.data_ptr()), info, POLICY_NO_LEVEL, vp(buf.data_ptr())) == 0
        torch.cuda.synchronize()
        M_cs = d_m.cpu().numpy()
        M_ours, levels, zp = cm.ilu0_host(a, ia, ja)
        assert zp == 0
        rel = np.abs(M_ours - M_cs) / np.maximum(np.abs(M_cs), 1e-300)
        assert rel.max() <= 1e-12, (nm, rel.max())
        cs.cusparseDestroyCsrilu02Info(info)
        cs.cusparseDestroyMatDescr(descr)

        # ---- sweeps: generic API SpSV on OUR factor (so only the sweep differs) ----
        CUSPARSE_INDEX_32I, CUSPARSE_INDEX_BASE_ZERO, CUDA_R_64F = 2, 0, 1
        FILL_MODE, DIAG_TYPE = 0, 1                       # cusparseSpMatAttribute_t
        FILL_LOWER, FILL_UPPER, DIAG_NON_UNIT, DIAG_UNIT = 0, 1, 0, 1
        d_mo = torch.from_numpy(M_ours).cuda()
        rhs = torch.from_numpy(np.random.default_rng(1).standard_normal(n)).cuda()
        s = cm.Solver(n)
        s.set_csr_host(a, ia, ja)
        s.analyze(cm.MODE_ILU0)
        for upper in (0, 1):
            mat, vx, vy, sp = vp(), vp(), vp(), vp()
            assert cs.cusparseCreateCsr(C.byref(mat), C.c_int64(n), C.c_int64(n), C.c_int64(nnz), vp(d_ia.data_ptr()), vp(d_ja.data_ptr()),
                                        vp(d_mo.data_ptr()), CUSPARSE_INDEX_32I, CUSPARSE_INDEX_32I, CUSPARSE_INDEX_BASE_ZERO, CUDA_R_64F) == 0
            fill = C.c_int(FILL_UPPER if upper else FILL_LOWER)
            diag = C.c_int(DIAG_NON_UNIT if upper else DIAG_UNIT)
            assert cs.cusparseSpMatSetAttribute(mat, FILL_MODE, C.byref(fill), C.c_size_t(4)) == 0
            assert cs.cusparseSpMatSetAttribute(mat, DIAG_TYPE, C.byref(diag), C.c_size_t(4)) == 0
            y = torch.zeros(n, dtype=torch.float64, device="cuda")
            assert cs.cusparseCreateDnVec(C.byref(vx), C.c_int64(n), vp(rhs.data_ptr()), CUDA_R_64F) == 0
            assert cs.cusparseCreateDnVec(C.byref(vy), C.c_int64(n), vp(y.data_ptr()), CUDA_R_64F) == 0
            assert cs.cusparseSpSV_createDescr(C.byref(sp)) == 0
            one = C.c_double(1.0)
            sz = C.c_size_t(0)
            assert cs.cusparseSpSV_bufferSize(h, 0, C.byref(one), mat, vx, vy, CUDA_R_64F, 0, sp, C.byref(sz)) == 0
            b2 = torch.empty(max(sz.value, 16), dtype=torch.uint8, device="cuda")
            assert cs.cusparseSpSV_analysis(h, 0, C.byref(one), mat, vx, vy, CUDA_R_64F, 0, sp, vp(b2.data_ptr())) == 0
            assert cs.cusparseSpSV_solve(h, 0, C.byref(one), mat, vx, vy, CUDA_R_64F, 0, sp) == 0
            ours = torch.zeros(n, dtype=torch.float64, device="cuda")
            s.sptrsv(upper, rhs.data_ptr(), ours.data_ptr())
            torch.cuda.synchronize()
            err = float((ours - y).abs().max() / y.abs().max())
            assert err <= 1e-11, (nm, upper, err)
            cs.cusparseSpSV_destroyDescr(sp); cs.cusparseDestroyDnVec(vx); cs.cusparseDestroyDnVec(vy); cs.cusparseDestroySpMat(mat)
        s.close()
    cs.cusparseDestroy(h)
