"""ctypes front-end of the CPU oracle (oracle/liboracle.so) and of the reference pieces built
into oracle/_ref.  TEST INFRASTRUCTURE ONLY — may be imported from tests/, from
__graft_entry__.smoke() and from bench.py's cpu_baseline / --impl reference legs, nowhere else.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF_MMIO = None
_REF_BICG = None

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)


class OrcStats(C.Structure):
    _fields_ = [("iterations", C.c_int), ("converged", C.c_int), ("breakdown", C.c_int),
                ("half_steps", C.c_int), ("nrm_r0", C.c_double), ("nrm_r", C.c_double)]


def build(ref=True):
    """(re)build liboracle.so and, if /root/reference exists, oracle/_ref."""
    subprocess.run(["make", "-s", "-C", _HERE], check=True)
    if ref and os.path.isdir(os.environ.get("CUDAMAT_REF", "/root/reference")):
        subprocess.run(["make", "-s", "-C", _HERE, "ref"], check=True)


def _dp(a):
    return a.ctypes.data_as(c_dp) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(c_ip) if a is not None else None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build(ref=False)
        L = C.CDLL(path)
        L.orc_dot.restype = C.c_double
        L.orc_dot.argtypes = [C.c_int64, c_dp, c_dp]
        L.orc_reduce_values.restype = C.c_double
        L.orc_reduce_values.argtypes = [C.c_int64, c_dp]
        L.orc_combine_tiles.restype = C.c_double
        L.orc_combine_tiles.argtypes = [C.c_int64, c_dp]
        L.orc_dot_tiles.argtypes = [C.c_int64, c_dp, c_dp, c_dp]
        L.orc_spmv.argtypes = [C.c_int, c_ip, c_ip, c_dp, c_dp, c_dp, c_dp]
        L.orc_ilu0.restype = C.c_int
        L.orc_ilu0.argtypes = [C.c_int, c_ip, c_ip, c_dp, c_dp]
        L.orc_sptrsv_lower_unit.argtypes = [C.c_int, c_ip, c_ip, c_dp, c_dp, c_dp]
        L.orc_sptrsv_upper.argtypes = [C.c_int, c_ip, c_ip, c_dp, c_dp, c_dp]
        L.orc_levels.restype = C.c_int
        L.orc_levels.argtypes = [C.c_int, c_ip, c_ip, C.c_int, c_ip]
        L.orc_bicgstab_unprec.restype = C.c_int
        L.orc_bicgstab_unprec.argtypes = [C.c_int, c_ip, c_ip, c_dp, c_dp, c_dp, c_dp, C.c_int,
                                          C.c_double, c_dp, C.POINTER(OrcStats), c_dp, C.c_int]
        L.orc_bicgstab_ilu0.restype = C.c_int
        L.orc_bicgstab_ilu0.argtypes = [C.c_int, c_ip, c_ip, c_dp, c_dp, C.c_int, C.c_double,
                                        c_dp, C.POINTER(OrcStats), c_dp, C.c_int]
        L.orc_bicgstab_ilu0_blocks.restype = C.c_int
        L.orc_bicgstab_ilu0_blocks.argtypes = [C.c_int, c_ip, c_ip, c_dp, C.c_int, C.POINTER(C.c_int64), c_dp, C.c_int, C.c_double,
                                               c_dp, C.POINTER(OrcStats), c_dp, C.c_int]
        L.orc_poisson3d.restype = C.c_int64
        L.orc_poisson3d.argtypes = [C.c_int, C.c_int64, C.c_int64, c_ip, c_ip, c_dp]
        L.orc_xtrue.argtypes = [C.c_uint64, C.c_int64, C.c_int64, c_dp]
        L.orc_random_dd.restype = C.c_int64
        L.orc_random_dd.argtypes = [C.c_int, C.c_uint64, c_ip, c_ip, c_dp]
        L.orc_glibc_rand_vector.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, c_dp]
        L.orc_to_dense_vector.argtypes = [C.c_int, c_dp, c_ip, c_dp]
        _LIB = L
    return _LIB


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def spmv(ia, ja, a, x, d=None):
    ia, ja, a, x = _i32(ia), _i32(ja), _f64(a), _f64(x)
    n = len(ia) - 1
    y = np.empty(n)
    d = _f64(d) if d is not None else None
    lib().orc_spmv(n, _ip(ia), _ip(ja), _dp(a), _dp(x), _dp(d), _dp(y))
    return y


def dot(a, b):
    a, b = _f64(a), _f64(b)
    return lib().orc_dot(len(a), _dp(a), _dp(b))


def dot_tiles(a, b):
    a, b = _f64(a), _f64(b)
    out = np.empty((len(a) + 2047) // 2048)
    lib().orc_dot_tiles(len(a), _dp(a), _dp(b), _dp(out))
    return out


def combine_tiles(tile):
    tile = _f64(tile)
    return lib().orc_combine_tiles(len(tile), _dp(tile))


def reduce_values(v):
    v = _f64(v)
    return lib().orc_reduce_values(len(v), _dp(v))


def ilu0(ia, ja, a):
    ia, ja, a = _i32(ia), _i32(ja), _f64(a)
    M = np.empty_like(a)
    st = lib().orc_ilu0(len(ia) - 1, _ip(ia), _ip(ja), _dp(a), _dp(M))
    return M, st


def sptrsv_lower_unit(ia, ja, M, rhs):
    ia, ja, M, rhs = _i32(ia), _i32(ja), _f64(M), _f64(rhs)
    out = np.empty(len(ia) - 1)
    lib().orc_sptrsv_lower_unit(len(ia) - 1, _ip(ia), _ip(ja), _dp(M), _dp(rhs), _dp(out))
    return out


def sptrsv_upper(ia, ja, M, rhs):
    ia, ja, M, rhs = _i32(ia), _i32(ja), _f64(M), _f64(rhs)
    out = np.empty(len(ia) - 1)
    lib().orc_sptrsv_upper(len(ia) - 1, _ip(ia), _ip(ja), _dp(M), _dp(rhs), _dp(out))
    return out


def levels(ia, ja, upper=False):
    ia, ja = _i32(ia), _i32(ja)
    lv = np.empty(len(ia) - 1, dtype=np.int32)
    nl = lib().orc_levels(len(ia) - 1, _ip(ia), _ip(ja), int(upper), _ip(lv))
    return lv, nl


def _stats(st, hist):
    return dict(iterations=st.iterations, converged=bool(st.converged), breakdown=st.breakdown,
                nrm_r0=st.nrm_r0, nrm_r=st.nrm_r, hist=hist[:st.half_steps].copy())


def bicgstab_unprec(ia, ja, a, b, d=None, x0=None, maxit=2000, tol=1e-6):
    ia, ja, a, b = _i32(ia), _i32(ja), _f64(a), _f64(b)
    n = len(ia) - 1
    d = _f64(d) if d is not None else None
    x0 = _f64(x0) if x0 is not None else None
    x = np.zeros(n)
    st = OrcStats()
    hist = np.zeros(maxit + 2)
    lib().orc_bicgstab_unprec(n, _ip(ia), _ip(ja), _dp(a), _dp(d), _dp(x0), _dp(b), maxit, tol,
                              _dp(x), C.byref(st), _dp(hist), len(hist))
    return x, _stats(st, hist)


def bicgstab_ilu0(ia, ja, a, b, maxit=2000, tol=1e-6):
    ia, ja, a, b = _i32(ia), _i32(ja), _f64(a), _f64(b)
    n = len(ia) - 1
    x = np.zeros(n)
    st = OrcStats()
    hist = np.zeros(2 * maxit + 2)
    lib().orc_bicgstab_ilu0(n, _ip(ia), _ip(ja), _dp(a), _dp(b), maxit, tol, _dp(x),
                            C.byref(st), _dp(hist), len(hist))
    return x, _stats(st, hist)


def bicgstab_ilu0_blocks(ia, ja, a, b, row_start, maxit=2000, tol=1e-6):
    """ILU0-BiCGSTAB with the block-Jacobi ILU(0) of the diagonal blocks [row_start[k], row_start[k+1]) (sharded handles)."""
    ia, ja, a, b = _i32(ia), _i32(ja), _f64(a), _f64(b)
    rs = np.ascontiguousarray(row_start, dtype=np.int64)
    n = len(ia) - 1
    assert rs[0] == 0 and rs[-1] == n and np.all(np.diff(rs) >= 0)
    x = np.zeros(n)
    st = OrcStats()
    hist = np.zeros(2 * maxit + 2)
    lib().orc_bicgstab_ilu0_blocks(n, _ip(ia), _ip(ja), _dp(a), len(rs) - 1, rs.ctypes.data_as(C.POINTER(C.c_int64)), _dp(b),
                                   maxit, tol, _dp(x), C.byref(st), _dp(hist), len(hist))
    return x, _stats(st, hist)


def poisson3d(N, row0=0, row1=None):
    n = N ** 3
    row1 = n if row1 is None else row1
    cnt = lib().orc_poisson3d(N, row0, row1, None, None, None)
    ia = np.empty(row1 - row0 + 1, dtype=np.int32)
    ja = np.empty(cnt, dtype=np.int32)
    a = np.empty(cnt)
    lib().orc_poisson3d(N, row0, row1, _ip(ia), _ip(ja), _dp(a))
    return ia, ja, a


def xtrue(seed, i0, cnt):
    out = np.empty(cnt)
    lib().orc_xtrue(seed, i0, cnt, _dp(out))
    return out


def random_dd(n, seed):
    ia = np.empty(n + 1, dtype=np.int32)
    nnz = lib().orc_random_dd(n, seed, _ip(ia), None, None)
    ja = np.empty(nnz, dtype=np.int32)
    a = np.empty(nnz)
    lib().orc_random_dd(n, seed, _ip(ia), _ip(ja), _dp(a))
    return ia, ja, a


def glibc_rand_vector(n, p_zero=0.2, vmin=1.0, vmax=5.0):
    out = np.empty(n)
    lib().orc_glibc_rand_vector(n, p_zero, vmin, vmax, _dp(out))
    return out


def to_dense_vector(n, A, IA):
    A, IA = _f64(A), _i32(IA)
    out = np.empty(n)
    lib().orc_to_dense_vector(n, _dp(A), _ip(IA), _dp(out))
    return out


# ---- reference pieces compiled into oracle/_ref (present only if built in this container) ----
def ref_available(which):
    return os.path.exists(os.path.join(_HERE, "_ref", {"mmio": "libref_mmio.so", "bicg": "libref_bicg.so",
                                                       "pbicgstab": "libref_pbicgstab.so"}[which]))


def ref_load_mm(path):
    """CSR exactly as the reference's loadMMSparseMatrix (mmio_wrapper.h:133-348) produces it."""
    global _REF_MMIO
    if _REF_MMIO is None:
        _REF_MMIO = C.CDLL(os.path.join(_HERE, "_ref", "libref_mmio.so"))
        _REF_MMIO.ref_loadMMSparseMatrix.argtypes = [C.c_char_p, c_ip, c_ip, c_ip, C.POINTER(c_dp),
                                                     C.POINTER(c_ip), C.POINTER(c_ip)]
        _REF_MMIO.ref_free.argtypes = [C.c_void_p]
    m, n, nnz = C.c_int(), C.c_int(), C.c_int()
    av, ai, aj = c_dp(), c_ip(), c_ip()
    rc = _REF_MMIO.ref_loadMMSparseMatrix(path.encode(), C.byref(m), C.byref(n), C.byref(nnz),
                                          C.byref(av), C.byref(ai), C.byref(aj))
    if rc != 0:
        raise RuntimeError("reference loadMMSparseMatrix failed for %s" % path)
    a = np.ctypeslib.as_array(av, (nnz.value,)).copy()
    ia = np.ctypeslib.as_array(ai, (m.value + 1,)).copy()
    ja = np.ctypeslib.as_array(aj, (nnz.value,)).copy()
    for p in (av, ai, aj):
        _REF_MMIO.ref_free(C.cast(p, C.c_void_p))
    return m.value, n.value, ia, ja, a


def ref_bicg(ia0, ja0, a, b, maxit=2000):
    """The reference's bicstab_omp BiCG() (bicstab.cpp:93-196), base-0 CSR. Returns (x, iters)."""
    global _REF_BICG
    if _REF_BICG is None:
        _REF_BICG = C.CDLL(os.path.join(_HERE, "_ref", "libref_bicg.so"))
        _REF_BICG.ref_bicg.argtypes = [C.c_int, C.c_int, c_dp, c_ip, c_ip, c_dp, c_dp, C.c_int, c_ip]
        _REF_BICG.ref_omp_threads.restype = C.c_int
    ia0, ja0, a, b = _i32(ia0), _i32(ja0), _f64(a), _f64(b)
    n = len(ia0) - 1
    x = np.zeros(n)
    it = C.c_int(0)
    _REF_BICG.ref_bicg(n, len(a), _dp(a), _ip(ja0), _ip(ia0), _dp(b), _dp(x), maxit, C.byref(it))
    return x, it.value


def ref_omp_threads(set_to=None):
    """OpenMP threads the reference's BiCG() will use; set_to = n forces omp_set_num_threads(n) first (torchrun exports
    OMP_NUM_THREADS=1 to its children, which would silently serialise the CPU baseline)."""
    ref_bicg(np.array([0, 1], dtype=np.int32), np.array([0], dtype=np.int32), np.array([1.0]), np.array([1.0]), 1)
    if set_to:
        _REF_BICG.ref_omp_set_threads(int(set_to))
    return _REF_BICG.ref_omp_threads()


# ---- the reference's own GPU solver: /root/reference/pbicgstab.cu compiled unmodified (oracle/Makefile) ----
_REF_GPU = None


def _ref_gpu():
    global _REF_GPU
    if _REF_GPU is None:
        L = C.CDLL(os.path.join(_HERE, "_ref", "libref_pbicgstab.so"))
        L.ref_bicgstab_lu_precond.argtypes = [C.c_int, C.c_int, c_dp, c_ip, c_ip, c_dp, C.c_int, C.c_double, c_dp, c_dp, C.c_char_p]
        L.ref_bicgstab_plain.argtypes = L.ref_bicgstab_lu_precond.argtypes
        L.ref_bicgstab_shifted.argtypes = [C.c_int, C.c_int, c_dp, c_ip, c_ip, c_dp, c_dp, c_dp, C.c_int, C.c_double, c_dp, c_dp, C.c_char_p]
        _REF_GPU = L
    return _REF_GPU


def _parse_ref_trace(path, mode):
    """The reference's debug lines (pbicgstab.cu:76,113,144 / :484,550) -> dict(nrm_r0, hist, iterations).
    `iterations` follows the loop counter i at exit: ILU0 loop: number of 'residual norm = ' lines (i is incremented
    after the second check only, :147-151); unpreconditioned loop: number of 'k = ' lines."""
    import re
    txt = open(path, errors="replace").read()
    out = {"trace": txt}
    if mode == "ilu0":
        m = re.search(r"init residual:norm\s+([-+0-9.eEinfa]+)", txt)
        out["nrm_r0"] = float(m.group(1)) if m else float("nan")
        half = [float(v) for v in re.findall(r"residual norm \(before precond\) = ([-+0-9.eEinfa]+)", txt)]
        full = [float(v) for v in re.findall(r"i = \d+, residual norm = ([-+0-9.eEinfa]+)", txt)]
        out["hist_half"], out["hist"], out["iterations"] = half, full, len(full)
    else:
        m = re.search(r"initial norm = ([-+0-9.eEinfa]+)", txt)
        out["nrm_r0"] = float(m.group(1)) if m else float("nan")
        full = [float(v) for v in re.findall(r"k = \d+, norm = ([-+0-9.eEinfa]+)", txt)]
        out["hist"], out["iterations"] = full, len(full)
    return out


def ref_gpu_bicgstab_lu_precond(ia, ja, a, b, maxit=2000, tol=1e-6, trace=True):
    """bicgstab_lu_precond of the reference (pbicgstab.cu:157), host arrays in the caller's index base.
    Returns (x, dtAlg, info); info carries the parsed debug trace when trace=True."""
    import tempfile
    ia, ja, a, b = _i32(ia).copy(), _i32(ja).copy(), _f64(a).copy(), _f64(b).copy()
    n = len(ia) - 1
    x = np.zeros(n)
    dt = C.c_double(0.0)
    with tempfile.NamedTemporaryFile(suffix=".trace") as tf:
        ok = _ref_gpu().ref_bicgstab_lu_precond(n, len(a), _dp(a), _ip(ia), _ip(ja), _dp(b), maxit, tol, _dp(x), C.byref(dt),
                                                tf.name.encode() if trace else None)
        info = _parse_ref_trace(tf.name, "ilu0") if trace else {}
    info["returned"] = bool(ok)
    return x, dt.value, info


def ref_gpu_bicgstab_shifted(ia, ja, a0, d, x0, b, maxit=2000, tol=1e-6, trace=True):
    """bicgstab(A0, d, x0, b) of the reference (pbicgstab.cu:926).  Valid for n <= 524288 only: mult_spec is launched
    with grid and block swapped (pbicgstab.cu:645,675,703)."""
    import tempfile
    ia, ja, a0, b = _i32(ia).copy(), _i32(ja).copy(), _f64(a0).copy(), _f64(b).copy()
    d, x0 = _f64(d).copy(), _f64(x0).copy()
    n = len(ia) - 1
    x = np.zeros(n)
    dt = C.c_double(0.0)
    with tempfile.NamedTemporaryFile(suffix=".trace") as tf:
        ok = _ref_gpu().ref_bicgstab_shifted(n, len(a0), _dp(a0), _ip(ia), _ip(ja), _dp(d), _dp(x0), _dp(b), maxit, tol, _dp(x),
                                             C.byref(dt), tf.name.encode() if trace else None)
        info = _parse_ref_trace(tf.name, "unprec") if trace else {}
    info["returned"] = bool(ok)
    return x, dt.value, info


def ref_gpu_bicgstab_plain(ia, ja, a, b, maxit=2000, tol=1e-6, trace=True):
    """bicgstab(A, b) of the reference (pbicgstab.cu:756): its initial residual is broken (pbicgstab.cu:469-478)."""
    import tempfile
    ia, ja, a, b = _i32(ia).copy(), _i32(ja).copy(), _f64(a).copy(), _f64(b).copy()
    n = len(ia) - 1
    x = np.zeros(n)
    dt = C.c_double(0.0)
    with tempfile.NamedTemporaryFile(suffix=".trace") as tf:
        ok = _ref_gpu().ref_bicgstab_plain(n, len(a), _dp(a), _ip(ia), _ip(ja), _dp(b), maxit, tol, _dp(x), C.byref(dt),
                                           tf.name.encode() if trace else None)
        info = _parse_ref_trace(tf.name, "unprec") if trace else {}
    info["returned"] = bool(ok)
    return x, dt.value, info
