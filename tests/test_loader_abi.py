"""CPU tests of the host side: the Matrix-Market loader against the reference loader's recorded output,
the C-ABI surface (every symbol of include/cudamat_b200.h is exported) and the no-GPU behaviour."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import GOLDEN, ROOT


@pytest.mark.parametrize("nm", ["mat3", "mat3_A0", "vec3", "vec3_d", "mat900", "mat10000"])
def test_loader_matches_reference_loader(cm, pin, nm):
    """ref_loader_csr.npz = what loadMMSparseMatrix (mmio_wrapper.h:133-348) returned for the reference's files"""
    m, n, ia, ja, a = cm.load_mm(os.path.join(GOLDEN, nm + ".mtx"))
    assert [m, n] == pin[nm + "_shape"].tolist()
    assert ia[0] == 1                                   # .mtx loads come out base-1
    assert np.array_equal(ia, pin[nm + "_ia"]) and np.array_equal(ja, pin[nm + "_ja"]) and np.array_equal(a, pin[nm + "_a"])


def test_loader_symmetric_expansion_and_csc(cm, tmp_path):
    p = tmp_path / "s.mtx"
    p.write_text("%%MatrixMarket matrix coordinate real skew-symmetric\n% c\n\n3 3 2\n2 1 5\n3 2 -7.5\n")
    m, n, ia, ja, a = cm.load_mm(str(p))
    assert (m, n) == (3, 3) and ia.tolist() == [1, 2, 4, 5] and ja.tolist() == [2, 1, 3, 2] and a.tolist() == [-5, 5, 7.5, -7.5]
    m, n, cp, ri, a2 = cm.load_mm(str(p), csr=False)
    assert cp.tolist() == [1, 2, 4, 5] and ri.tolist() == [2, 1, 3, 2] and a2.tolist() == [5, -5, -7.5, 7.5]
    # integer field is accepted, base-0 files are detected
    p.write_text("%%MatrixMarket matrix coordinate integer general\n2 2 2\n0 0 3\n1 1 4\n")
    m, n, ia, ja, a = cm.load_mm(str(p))
    assert ia.tolist() == [0, 1, 2] and ja.tolist() == [0, 1] and a.tolist() == [3, 4]


@pytest.mark.parametrize("text,why", [
    ("%%MatrixMarket matrix array real general\n2 2\n1\n2\n3\n4\n", "array"),
    ("%%MatrixMarket matrix coordinate pattern general\n2 2 1\n1 1\n", "pattern"),
    ("%%MatrixMarket matrix coordinate complex general\n2 2 1\n1 1 1 0\n", "complex"),
    ("%%MatrixMarket matrix coordinate real general\n2 2 2\n1 1 1\n1 1 2\n", "duplicate"),
    ("%%MatrixMarket matrix coordinate real general\n2 2 2\n0 0 1\n2 2 2\n", "base-0 and base-1"),
    ("%%MatrixMarket matrix coordinate real general\n2 2 3\n1 1 1\n", "truncated"),
    ("hello\n", "banner"),
])
def test_loader_rejects(cm, tmp_path, text, why):
    p = tmp_path / "bad.mtx"
    p.write_text(text)
    with pytest.raises(cm.CudamatError):
        cm.load_mm(str(p))
    with pytest.raises(cm.CudamatError):
        cm.load_mm(str(tmp_path / "missing.mtx"))


def test_writer_round_trip_and_large_parallel_parse(cm, O, pin, tmp_path):
    """cudamat_write_mm / cudamat_write_mm_vector (mm_write_mtx_crd, mmio.c:405) round-trip exactly through the loader,
    general and symmetric-lower; a > 1 MiB file goes through the multi-threaded parser and must agree with the oracle's CSR."""
    ia, ja, a = csr = (pin["mat900_ia"], pin["mat900_ja"], pin["mat900_a"])
    for sym in (False, True):
        p = str(tmp_path / ("w%d.mtx" % sym))
        cm.write_mm(p, 900, 900, ia, ja, a, symmetric=sym, comment="round trip")
        m, n, ia2, ja2, a2 = cm.load_mm(p)
        assert (m, n) == (900, 900) and np.array_equal(ia2, ia) and np.array_equal(ja2, ja) and np.array_equal(a2, a)
    # base-0 CSR with awkward values
    pia, pja, pa = O.poisson3d(40)                       # 64000 rows, 438400 entries: ~9 MB of text -> parallel parse
    rng = np.random.default_rng(3)
    pa = pa * rng.standard_normal(len(pa)) * 10.0 ** rng.integers(-30, 30, len(pa))
    p = str(tmp_path / "big.mtx")
    cm.write_mm(p, 64000, 64000, pia, pja, pa)
    assert os.path.getsize(p) > (1 << 20)
    m, n, ia2, ja2, a2 = cm.load_mm(p)
    assert np.array_equal(ia2 - 1, pia) and np.array_equal(ja2 - 1, pja) and np.array_equal(a2, pa)
    x = rng.standard_normal(777); x[5] = 0.0; x[776] = 0.0
    p = str(tmp_path / "v.mtx")
    cm.write_mm_vector(p, x)
    m, n, vi, vj, va = cm.load_mm(p)
    assert (m, n) == (777, 1) and np.array_equal(cm.to_dense_vector(777, va, vi), x)
    with pytest.raises(cm.CudamatError):
        cm.write_mm(str(tmp_path / "nodir" / "x.mtx"), 900, 900, ia, ja, a)


def test_loader_agrees_with_reference_loader_on_generated_files(cm, O, tmp_path):
    """property test against the REFERENCE's loader (oracle/_ref, only where the reference checkout was compiled):
    random real / integer files, general / symmetric / skew-symmetric / hermitian, base-1."""
    if not O.ref_available("mmio"):
        pytest.skip("oracle/_ref/libref_mmio.so not built (needs /root/reference)")
    rng = np.random.default_rng(11)
    for field in ("real", "integer"):
        for sym in ("general", "symmetric", "skew-symmetric", "hermitian"):
            n = int(rng.integers(5, 40))
            ents = {}
            for _ in range(4 * n):
                i, j = int(rng.integers(1, n + 1)), int(rng.integers(1, n + 1))
                if sym != "general" and j > i:
                    i, j = j, i
                if sym == "skew-symmetric" and i == j:
                    continue
                ents[(i, j)] = int(rng.integers(-9, 10)) if field == "integer" else float(rng.standard_normal())
            ents[(n, n if sym != "skew-symmetric" else 1)] = 3 if field == "integer" else 3.5      # makes the file base-1
            p = tmp_path / ("%s_%s.mtx" % (field, sym))
            with open(p, "w") as f:
                f.write("%%%%MatrixMarket matrix coordinate %s %s\n%% generated\n%d %d %d\n" % (field, sym, n, n, len(ents)))
                for (i, j), v in ents.items():
                    f.write(("%d %d %d\n" if field == "integer" else "%d %d %.17g\n") % (i, j, v))
            if sym == "hermitian" and field == "real":   # mm_is_valid (mmio.c:96) rejects real hermitian: both loaders must fail
                with pytest.raises(cm.CudamatError):
                    cm.load_mm(str(p))
                with pytest.raises(RuntimeError):
                    O.ref_load_mm(str(p))
                continue
            got = cm.load_mm(str(p))
            ref = O.ref_load_mm(str(p))
            assert got[0] == ref[0] and got[1] == ref[1], (field, sym)
            for g, r in zip(got[2:], ref[2:]):
                assert np.array_equal(g, r), (field, sym)


def test_to_dense_vector(cm, O, pin):
    for nm in ("vec3", "vec3_d"):
        got = cm.to_dense_vector(3, pin[nm + "_a"], pin[nm + "_ia"])
        assert np.array_equal(got, O.to_dense_vector(3, pin[nm + "_a"], pin[nm + "_ia"]))


def test_abi_exports_every_declared_symbol(cm):
    hdr = open(os.path.join(ROOT, "include", "cudamat_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(cudamat_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 25
    lib = ctypes.CDLL(cm.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), "missing export " + name
    assert sorted(cm.EXPORTS) == declared
    assert lib.cudamat_abi_version() == 2


def test_poisson_nnz_closed_form(cm, O):
    for N in (1, 2, 5, 16):
        n = N ** 3
        ia, _, _ = O.poisson3d(N)
        assert cm.poisson3d_nnz(N) == ia[-1]
        for r0, r1 in ((0, n // 3), (n // 3, n), (7 % n, max(7 % n, n - 3))):
            assert cm.poisson3d_nnz(N, r0, r1) == ia[r1] - ia[r0]
    assert cm.poisson3d_nnz(256) == 117047296 and cm.poisson3d_nnz(512) == 937951232


def test_no_cpu_fallback(cm):
    """without a device every compute entry point must fail loudly (never route through a CPU path)"""
    if cm.device_count() > 0:
        pytest.skip("a GPU is present")
    one = np.array([1.0]); ia = np.array([0, 1], dtype=np.int32); ja = np.array([0], dtype=np.int32)
    for fn in (cm.bicgstab, cm.bicgstab_lu_precond):
        with pytest.raises(cm.CudamatError) as e:
            fn(one, ia, ja, one)
        assert e.value.code == cm.E_NO_DEVICE
    with pytest.raises(cm.CudamatError):
        cm.Solver(10)
    with pytest.raises(cm.CudamatError):
        cm.ilu0_host(one, ia, ja)
