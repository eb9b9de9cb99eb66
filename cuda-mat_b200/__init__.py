"""cuda_mat_b200 — Python (ctypes) binding of libcudamat_b200.so, the B200-native BiCGSTAB path.

This package is a thin test/bench harness over the C ABI declared in include/cudamat_b200.h; the
product is the shared library (hand-written sm_100a kernels + C++ host code).  The directory name
contains a hyphen, so load it with ``load_package()`` from ``__graft_entry__`` / ``tests/conftest``
(importlib under the module name ``cuda_mat_b200``).

There is NO CPU fallback: if the library is missing this module raises at import time, and every
compute entry point returns CUDAMAT_E_NO_DEVICE without a GPU.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CUDAMAT_LIB") or os.path.join(_HERE, "libcudamat_b200.so")   # CUDAMAT_LIB: A/B tuning builds
ROOT = os.path.dirname(_HERE)

MODE_PLAIN, MODE_SHIFTED, MODE_ILU0 = 0, 1, 2
SPMV_AUTO, SPMV_ROWLANE, SPMV_STAGED, SPMV_PATTERN, SPMV_CLASS, SPMV_TILED, SPMV_MARCH, SPMV_STREAM = 0, 1, 2, 3, 4, 5, 6, 7
E_NO_DEVICE = -2

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)


class Stats(C.Structure):
    _fields_ = [("iterations", C.c_int), ("converged", C.c_int), ("breakdown", C.c_int),
                ("half_steps", C.c_int), ("nrm_r0", C.c_double), ("nrm_r", C.c_double),
                ("t_h2d", C.c_double), ("t_analysis", C.c_double), ("t_ilu0", C.c_double),
                ("t_loop", C.c_double), ("t_d2h", C.c_double), ("levels_l", C.c_int),
                ("levels_u", C.c_int), ("spmv_variant", C.c_int), ("zero_pivot", C.c_int),
                ("kernel_launches", C.c_int64), ("t_spmv", C.c_double), ("n_spmv", C.c_int),
                ("graph_replay", C.c_int), ("t_kernel", C.c_double * 4), ("n_kernel", C.c_int * 4), ("fused", C.c_int)]

    def as_dict(self):
        return {k: (list(getattr(self, k)) if k in ("t_kernel", "n_kernel") else getattr(self, k)) for k, _ in self._fields_}


class CudamatError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("cudamat error %d: %s" % (code, msg))
        self.code = code


def build(verbose=False):
    """Compile libcudamat_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
    subprocess.run(["make", "-j4", "-C", _HERE] + ([] if verbose else ["-s"]), check=True)


EXPORTS = [
    "cudamat_abi_version", "cudamat_last_error", "cudamat_device_count", "cudamat_bicgstab_host",
    "cudamat_ilu0_host", "cudamat_create", "cudamat_destroy", "cudamat_set_option",
    "cudamat_set_csr_host", "cudamat_set_csr_device", "cudamat_analyze", "cudamat_solve_device",
    "cudamat_get_history", "cudamat_spmv_device", "cudamat_dot_device", "cudamat_get_ilu0_host",
    "cudamat_sptrsv_device", "cudamat_sweep_blocks", "cudamat_comm_p2p_enabled", "cudamat_write_mm", "cudamat_write_mm_vector", "cudamat_comm_unique_id", "cudamat_comm_init", "cudamat_partition_rows",
    "cudamat_halo_plan_host", "cudamat_tiled_plan_host", "cudamat_march_plan_host",
    "cudamat_gen_poisson3d_device", "cudamat_poisson3d_nnz", "cudamat_gen_xtrue_device",
    "cudamat_gen_random_dd_device", "cudamat_load_mm", "cudamat_free",
]

if not os.path.exists(LIB_PATH):
    raise ImportError("libcudamat_b200.so is not built (run __graft_entry__.build() or make -C cuda-mat_b200); "
                      "there is no CPU fallback")
lib = C.CDLL(LIB_PATH)

lib.cudamat_last_error.restype = C.c_char_p
lib.cudamat_tiled_plan_host.argtypes = [C.c_int, c_ip, c_ip, c_dp, C.POINTER(C.c_uint), C.c_int, C.c_int, c_ip, c_ip, c_ip, c_ip, c_ip,
                                        C.POINTER(C.c_ulonglong), c_ip, c_ip, c_dp, C.POINTER(C.c_ubyte), C.POINTER(C.c_longlong)]
lib.cudamat_bicgstab_host.argtypes = [C.c_int, C.c_int, C.c_int, c_dp, c_ip, c_ip, c_dp, c_dp, c_dp, C.c_int,
                                      C.c_double, C.c_int, c_dp, c_dp, C.POINTER(Stats)]
lib.cudamat_ilu0_host.argtypes = [C.c_int, C.c_int, c_dp, c_ip, c_ip, c_dp, c_ip, c_ip]
lib.cudamat_create.argtypes = [C.POINTER(C.c_void_p), C.c_int64, C.c_int64, C.c_int64, C.c_void_p]
lib.cudamat_destroy.argtypes = [C.c_void_p]
lib.cudamat_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_int64]
lib.cudamat_set_csr_host.argtypes = [C.c_void_p, C.c_int, c_dp, c_ip, c_ip]
lib.cudamat_set_csr_device.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
lib.cudamat_analyze.argtypes = [C.c_void_p, C.c_int, C.POINTER(Stats)]
lib.cudamat_solve_device.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_int, C.c_double, C.POINTER(Stats)]
lib.cudamat_get_history.argtypes = [C.c_void_p, c_dp, C.c_int]
lib.cudamat_spmv_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
lib.cudamat_dot_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, c_dp]
lib.cudamat_get_ilu0_host.argtypes = [C.c_void_p, c_dp]
lib.cudamat_sptrsv_device.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
lib.cudamat_sweep_blocks.argtypes = [C.c_void_p]
lib.cudamat_partition_rows.argtypes = [C.c_int64, C.c_int, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
lib.cudamat_halo_plan_host.argtypes = [C.c_int64, C.c_int64, C.c_int64, c_ip, C.c_int, C.POINTER(C.c_int64), c_ip,
                                       C.POINTER(c_ip), c_ip]
lib.cudamat_comm_unique_id.argtypes = [C.c_void_p]
lib.cudamat_comm_init.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
lib.cudamat_comm_p2p_enabled.argtypes = [C.c_void_p]
lib.cudamat_gen_poisson3d_device.argtypes = [C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
lib.cudamat_poisson3d_nnz.restype = C.c_int64
lib.cudamat_poisson3d_nnz.argtypes = [C.c_int, C.c_int64, C.c_int64]
lib.cudamat_gen_xtrue_device.argtypes = [C.c_uint64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]
lib.cudamat_gen_random_dd_device.argtypes = [C.c_int, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.POINTER(C.c_int64), C.c_void_p]
lib.cudamat_load_mm.argtypes = [C.c_char_p, C.c_int, c_ip, c_ip, c_ip, C.POINTER(c_dp), C.POINTER(c_ip), C.POINTER(c_ip)]
lib.cudamat_free.argtypes = [C.c_void_p]
lib.cudamat_write_mm.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, c_dp, c_ip, c_ip, C.c_int, C.c_char_p]
lib.cudamat_write_mm_vector.argtypes = [C.c_char_p, C.c_int, c_dp, C.c_char_p]


def last_error():
    return lib.cudamat_last_error().decode(errors="replace")


def _check(rc):
    if rc != 0:
        raise CudamatError(rc, last_error())


def device_count():
    return lib.cudamat_device_count()


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _dp(a):
    return None if a is None else a.ctypes.data_as(c_dp)


def _ip(a):
    return a.ctypes.data_as(c_ip)


# ------------------------------------------------------------------------------------------------
# host-pointer API: mirrors of the reference's three entry points (pbicgstab.h:113,116,119-120)
# ------------------------------------------------------------------------------------------------
def _solve_host(mode, A, iA, jA, b, d=None, x0=None, maxit=2000, tol=1e-6, debug=False):
    A, iA, jA, b, d, x0 = _f64(A), _i32(iA), _i32(jA), _f64(b), _f64(d), _f64(x0)
    n = len(iA) - 1
    x = np.zeros(n)
    dt = C.c_double(0.0)
    st = Stats()
    _check(lib.cudamat_bicgstab_host(mode, n, len(A), _dp(A), _ip(iA), _ip(jA), _dp(d), _dp(x0), _dp(b),
                                     maxit, tol, int(debug), _dp(x), C.byref(dt), C.byref(st)))
    return x, dt.value, st.as_dict()


def bicgstab(A, iA, jA, b, maxit=2000, tol=1e-6, debug=False):
    """solve Ax = b, no preconditioner (pbicgstab.h:113; x0 = ones)."""
    return _solve_host(MODE_PLAIN, A, iA, jA, b, maxit=maxit, tol=tol, debug=debug)


def bicgstab_shifted(A0, iA0, jA0, d, x0, b, maxit=2000, tol=1e-6, debug=False):
    """solve (A0 + I*d)x = b, no preconditioner (pbicgstab.h:116)."""
    return _solve_host(MODE_SHIFTED, A0, iA0, jA0, b, d=d, x0=x0, maxit=maxit, tol=tol, debug=debug)


def bicgstab_lu_precond(A, iA, jA, b, maxit=2000, tol=1e-6, debug=False):
    """solve Ax = b with the ILU0 right preconditioner (pbicgstab.h:119-120; x0 = ones)."""
    return _solve_host(MODE_ILU0, A, iA, jA, b, maxit=maxit, tol=tol, debug=debug)


def ilu0_host(A, iA, jA):
    A, iA, jA = _f64(A), _i32(iA), _i32(jA)
    M = np.empty_like(A)
    lv = (C.c_int * 2)()
    zp = C.c_int(0)
    _check(lib.cudamat_ilu0_host(len(iA) - 1, len(A), _dp(A), _ip(iA), _ip(jA), _dp(M), lv, C.byref(zp)))
    return M, (lv[0], lv[1]), zp.value


def load_mm(path, csr=True):
    """loadMMSparseMatrix(filename, 'd', csr, ...) (mmio_wrapper.h:133): returns (m, n, ptr, ind, val)."""
    m, n, nnz = C.c_int(), C.c_int(), C.c_int()
    av, ai, aj = c_dp(), c_ip(), c_ip()
    _check(lib.cudamat_load_mm(path.encode(), int(csr), C.byref(m), C.byref(n), C.byref(nnz),
                               C.byref(av), C.byref(ai), C.byref(aj)))
    major = m.value if csr else n.value
    ptr_p, ind_p = (ai, aj) if csr else (aj, ai)
    val = np.ctypeslib.as_array(av, (max(nnz.value, 1),))[:nnz.value].copy()
    ptr = np.ctypeslib.as_array(ptr_p, (major + 1,)).copy()
    ind = np.ctypeslib.as_array(ind_p, (max(nnz.value, 1),))[:nnz.value].copy()
    for p in (av, ai, aj):
        lib.cudamat_free(C.cast(p, C.c_void_p))
    return m.value, n.value, ptr, ind, val


def write_mm(path, m, n, ptr, ind, val, symmetric=False, comment=None):
    """CSR -> Matrix Market coordinate file (replaces mm_write_mtx_crd, mmio.c:405-445)."""
    ptr, ind, val = _i32(ptr), _i32(ind), _f64(val)
    _check(lib.cudamat_write_mm(path.encode(), m, n, len(val), _dp(val), _ip(ptr), _ip(ind), int(symmetric),
                                comment.encode() if comment else None))


def write_mm_vector(path, x, comment=None):
    """dense vector -> n x 1 coordinate file (the reference's -V input, example.cpp:310-336)."""
    x = _f64(x)
    _check(lib.cudamat_write_mm_vector(path.encode(), len(x), _dp(x), comment.encode() if comment else None))


def to_dense_vector(n, A, IA):
    """toDenseVector (pbicgstab.cu:1101-1115): m x 1 CSR column vector -> dense."""
    out = np.zeros(n)
    s, cnt = IA[0], 0
    for i in range(n):
        if IA[i + 1] - s > 0:
            out[i] = A[cnt]
            cnt += 1
            s = IA[i + 1]
    return out


# ------------------------------------------------------------------------------------------------
# handle API on raw device pointers (torch tensors: pass t.data_ptr())
# ------------------------------------------------------------------------------------------------
class Solver:
    def __init__(self, n_global, row0=0, row1=None, stream=0):
        row1 = n_global if row1 is None else row1
        self.h = C.c_void_p()
        self.n = row1 - row0
        _check(lib.cudamat_create(C.byref(self.h), n_global, row0, row1, C.c_void_p(stream)))
        self._keep = []

    def close(self):
        if self.h:
            lib.cudamat_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, key, value):
        _check(lib.cudamat_set_option(self.h, key.encode(), int(value)))

    def set_csr_host(self, A, iA, jA):
        A, iA, jA = _f64(A), _i32(iA), _i32(jA)
        _check(lib.cudamat_set_csr_host(self.h, len(A), _dp(A), _ip(iA), _ip(jA)))

    def set_csr_device(self, nnz, dA, dIA, dJA, keep=None):
        self._keep.append(keep)
        _check(lib.cudamat_set_csr_device(self.h, nnz, dA, dIA, dJA))

    def analyze(self, mode):
        st = Stats()
        _check(lib.cudamat_analyze(self.h, mode, C.byref(st)))
        return st.as_dict()

    def solve(self, mode, d_b, d_x, d_x0=None, d_d=None, maxit=2000, tol=1e-6):
        st = Stats()
        _check(lib.cudamat_solve_device(self.h, mode, d_b, d_x0, d_d, d_x, maxit, tol, C.byref(st)))
        return st.as_dict()

    def history(self, cap=100000):
        buf = np.zeros(cap)
        m = lib.cudamat_get_history(self.h, _dp(buf), cap)
        return buf[:m].copy()

    def spmv(self, d_x, d_y, d_d=None, variant=SPMV_AUTO):
        _check(lib.cudamat_spmv_device(self.h, d_x, d_d, d_y, variant))

    def dot(self, d_a, d_b):
        r = C.c_double(0.0)
        _check(lib.cudamat_dot_device(self.h, d_a, d_b, C.byref(r)))
        return r.value

    def ilu0_values(self, nnz):
        M = np.empty(nnz)
        _check(lib.cudamat_get_ilu0_host(self.h, _dp(M)))
        return M

    def sptrsv(self, upper, d_rhs, d_out):
        _check(lib.cudamat_sptrsv_device(self.h, int(upper), d_rhs, d_out))

    def sweep_blocks(self):
        return int(lib.cudamat_sweep_blocks(self.h))


def partition_rows(n_global, world, rank):
    """contiguous row shard [row0,row1) of `rank`, aligned to the reduction group / tile"""
    r0, r1 = C.c_int64(0), C.c_int64(0)
    _check(lib.cudamat_partition_rows(n_global, world, rank, C.byref(r0), C.byref(r1)))
    return r0.value, r1.value


def halo_plan_host(row0, row1, ja_global, row_starts):
    """(halo_cols, recv_cnt): sorted unique off-shard columns and how many each owner rank holds"""
    ja_global = _i32(ja_global)
    world = len(row_starts) - 1
    rs = (C.c_int64 * (world + 1))(*[int(v) for v in row_starts])
    nh = C.c_int(0)
    hp = c_ip()
    rc = (C.c_int * world)()
    _check(lib.cudamat_halo_plan_host(row0, row1, len(ja_global), _ip(ja_global), world, rs, C.byref(nh), C.byref(hp), rc))
    halo = np.ctypeslib.as_array(hp, (max(nh.value, 1),))[:nh.value].copy()
    lib.cudamat_free(C.cast(hp, C.c_void_p))
    return halo, np.array(list(rc), dtype=np.int64)


def march_plan_host(sup_off, sup_val, n):
    """Host planner of the MARCH SpMV variant: None, or dict(D, H, S, P, dz, loff) for the superset pattern sup_off."""
    so = _i32(sup_off)
    sv = _f64(sup_val if sup_val is not None else np.ones(len(so)))
    ok, D, H, S, P = C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0)
    dz, lo = (C.c_int * 8)(), (C.c_int * 8)()
    lib.cudamat_march_plan_host.argtypes = [C.c_int, c_ip, c_dp, C.c_longlong, c_ip, c_ip, c_ip, c_ip, c_ip, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    _check(lib.cudamat_march_plan_host(len(so), _ip(so), _dp(sv), int(n), C.byref(ok), C.byref(D), C.byref(H), C.byref(S), C.byref(P), dz, lo))
    if not ok.value:
        return None
    return {"D": D.value, "H": H.value, "S": S.value, "P": P.value, "dz": list(dz)[:len(so)], "loff": list(lo)[:len(so)]}


def tiled_plan_host(lens, offs, vals, hist, n, with_vals=True):
    """Host planner of the TILED SpMV variant (pure host code).  lens[c], offs[c][q] (column offsets in storage order),
    vals[c][q] or None, hist[c] rows per class, n rows.  Returns None when there is no plan, else a dict with the staged
    windows, the shared-memory index of every entry, the classes that fit and the superset pattern (sup_len 0: none)."""
    ncls = len(lens)
    L = _i32(lens)
    O = np.zeros((ncls, 16), dtype=np.int32)
    V = np.zeros((ncls, 16), dtype=np.float64)
    for c in range(ncls):
        O[c, :len(offs[c])] = offs[c]
        if vals is not None:
            V[c, :len(vals[c])] = vals[c]
    H = np.ascontiguousarray(hist, dtype=np.uint32)
    nseg, sup_len = C.c_int(0), C.c_int(0)
    seg_lo, seg_len, seg_base = (C.c_int * 4)(), (C.c_int * 4)(), (C.c_int * 4)()
    disp = np.zeros((ncls, 16), dtype=np.int32)
    ok = C.c_ulonglong(0)
    sup_boff = (C.c_int * 8)()
    sup_val = (C.c_double * 8)()
    cmask = (C.c_ubyte * 64)()
    smem = C.c_longlong(0)
    _check(lib.cudamat_tiled_plan_host(ncls, _ip(L), _ip(O), _dp(V) if vals is not None else None,
                                       H.ctypes.data_as(C.POINTER(C.c_uint)), int(n), 1 if with_vals else 0, C.byref(nseg), seg_lo, seg_len,
                                       seg_base, _ip(disp), C.byref(ok), C.byref(sup_len), sup_boff, sup_val, cmask, C.byref(smem)))
    if nseg.value == 0:
        return None
    k = nseg.value
    return {"windows": [(seg_lo[g], seg_len[g], seg_base[g]) for g in range(k)], "disp": disp, "ok_mask": ok.value,
            "sup_len": sup_len.value, "sup_boff": list(sup_boff)[:sup_len.value], "sup_val": list(sup_val)[:sup_len.value],
            "class_mask": list(cmask)[:ncls], "smem_bytes": smem.value}


class Comm:
    """NCCL communicator of a sharded Solver; the unique id travels through torch.distributed."""

    @staticmethod
    def unique_id():
        raw = (C.c_ubyte * 128)()
        _check(lib.cudamat_comm_unique_id(raw))
        return bytes(raw)

    @staticmethod
    def init(solver, uid, rank, world):
        raw = (C.c_ubyte * 128)(*uid)
        _check(lib.cudamat_comm_init(solver.h, raw, rank, world))

    @staticmethod
    def p2p_enabled(solver):
        return bool(lib.cudamat_comm_p2p_enabled(solver.h))


def gen_poisson3d_device(N, row0, row1, d_ia, d_ja, d_a, stream=0):
    _check(lib.cudamat_gen_poisson3d_device(N, row0, row1, d_ia, d_ja, d_a, C.c_void_p(stream)))


def poisson3d_nnz(N, row0=0, row1=None):
    return lib.cudamat_poisson3d_nnz(N, row0, N ** 3 if row1 is None else row1)


def gen_xtrue_device(seed, i0, cnt, d_out, stream=0):
    _check(lib.cudamat_gen_xtrue_device(seed, i0, cnt, d_out, C.c_void_p(stream)))


def gen_random_dd_device(n, seed, d_ia, d_ja=None, d_a=None, stream=0):
    nnz = C.c_int64(0)
    _check(lib.cudamat_gen_random_dd_device(n, seed, d_ia, d_ja, d_a, C.byref(nnz), C.c_void_p(stream)))
    return nnz.value
