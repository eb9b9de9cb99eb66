"""Parity against the REFERENCE ITSELF, run on the same GPU (pytest -m gpu).

oracle/_ref/libref_pbicgstab.so is /root/reference/pbicgstab.cu compiled unmodified (oracle/Makefile; the six legacy
cuSPARSE names it calls are mapped onto their successors inside cuSPARSE by the test-only header
oracle/ref_shims/refgpu/legacy_cusparse.h).  Both sides get the same host arrays:

    reference  bicgstab_lu_precond  pbicgstab.cu:157   vs   cudamat_bicgstab_host(MODE_ILU0)
    reference  bicgstab(A0,d,x0,b)  pbicgstab.cu:926   vs   cudamat_bicgstab_host(MODE_SHIFTED)

Bars (BASELINE.json north_star): iteration count within +-2, ||x - x_ref|| <= 1e-8 ||x_ref||, final relative residual
<= tol.  The reference's iteration count is read from its own debug trace.

Measured on B200 (profiles/r2_reference_parity.json): on the reference's own fixtures in ILU0 mode (its only live path,
example.cpp:352) the counts agree to +-1 and x to <= 2e-7 (1e-6 cases) / 6e-11 (1e-10 cases).  On erratically converging
systems (unpreconditioned mat10000, Poisson >= 64^3) the REFERENCE DOES NOT REPRODUCE ITS OWN COUNT within +-2 when b is
perturbed in the last bit (BiCGSTAB is chaotic w.r.t. rounding, and the reference's summation orders live
inside cuBLAS/cuSPARSE).  The assertable form of the bar is therefore relative to the reference's own reproducibility:
every case is run 1 + NPERT times on both sides, b perturbed by <= 1 ulp per entry, and
    * our count range must meet [min_ref - 2, max_ref + 2];
    * ||x - x_ref|| <= 1e-8 ||x_ref||  OR  <= 4 x the largest distance between two reference runs (two converged solves of an
      ill-conditioned system differ by up to cond(A) * tol whoever computes them);
    * our true relative residual <= tol.
"""
import json
import os

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref(O, torch_cuda):
    if not O.ref_available("pbicgstab"):
        pytest.skip("oracle/_ref/libref_pbicgstab.so not built (needs /root/reference + nvcc at build time)")
    return O


def _relres(O, ia0, ja0, a, b, x, d=None):
    r = b - O.spmv(ia0, ja0, a, x, d=d)
    return float(np.linalg.norm(r))


TABLE = []


def _record(row):
    TABLE.append(row)
    out = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "ref_parity.json"), "w") as f:
            json.dump(TABLE, f, indent=1)
    except OSError:
        pass
    print("ref-parity", json.dumps(row))


NPERT = 3


def _perturbed(b, k):
    """b with every entry moved by at most one unit in the last place (deterministic in k)"""
    if k == 0:
        return b
    rng = np.random.default_rng(1000 + k)
    return b * (1.0 + rng.integers(-1, 2, size=len(b)) * 2.0 ** -52)


def _band(run_ref, run_ours, b):
    """run both sides on b and on NPERT last-bit perturbations; returns the unperturbed results and the spreads"""
    ref = [run_ref(_perturbed(b, k)) for k in range(NPERT + 1)]
    ours = [run_ours(_perturbed(b, k)) for k in range(NPERT + 1)]
    it_ref = [r[2]["iterations"] for r in ref]
    it_ours = [r[2]["iterations"] for r in ours]
    xr = ref[0][0]
    nx = np.linalg.norm(xr)
    ref_self = max(float(np.linalg.norm(r[0] - xr) / nx) for r in ref[1:])
    return ref[0], ours[0], it_ref, it_ours, ref_self


def _judge(name, tol, it_ref, it_ours, xerr, ref_self, rr_ours, bad):
    if max(it_ours) < min(it_ref) - 2 or min(it_ours) > max(it_ref) + 2:
        bad.append((name, "iterations", it_ref, it_ours))
    if rr_ours > tol * 1.0000001:
        bad.append((name, "relres", rr_ours))
    if xerr > max(1e-8, 4.0 * ref_self):
        bad.append((name, "xerr", xerr, ref_self))


def _poisson_case(O, N):
    ia, ja, a = O.poisson3d(N)
    xt = O.xtrue(1234, 0, N ** 3)
    return ia, ja, a, O.spmv(ia, ja, a, xt)


def _ilu0_cases(O, pin):
    cases = []
    for nm in ("mat900", "mat10000"):
        ia, ja, a = pin[nm + "_ia"], pin[nm + "_ja"], pin[nm + "_a"]      # base-1, exactly what the reference loader yields
        n = len(ia) - 1
        for bname, b in (("ones", np.ones(n)), ("glibc_rand", O.glibc_rand_vector(n))):
            for tol in (1e-6, 1e-10):
                cases.append(("%s/%s/%g" % (nm, bname, tol), ia, ja, a, b, tol))
    for N in (32, 64, 128):
        ia, ja, a, b = _poisson_case(O, N)
        cases.append(("poisson%d/Axtrue/1e-10" % N, ia, ja, a, b, 1e-10))
    return cases


def test_ilu0_vs_reference_bicgstab_lu_precond(cm, ref, pin):
    O = ref
    bad = []
    for name, ia, ja, a, b, tol in _ilu0_cases(O, pin):
        base = int(ia[0])
        (xr, dtr, info), (x, dt, st), it_ref, it_ours, ref_self = _band(
            lambda bb: O.ref_gpu_bicgstab_lu_precond(ia, ja, a, bb, maxit=2000, tol=tol),
            lambda bb: cm.bicgstab_lu_precond(a, ia, ja, bb, maxit=2000, tol=tol), b)
        ia0, ja0 = ia - base, ja - base
        nrm0 = _relres(O, ia0, ja0, a, b, np.ones(len(b)))
        rr_ours = _relres(O, ia0, ja0, a, b, x) / nrm0
        rr_ref = _relres(O, ia0, ja0, a, b, xr) / nrm0
        xerr = float(np.linalg.norm(x - xr) / np.linalg.norm(xr))
        row = dict(case=name, mode="ilu0", n=len(b), tol=tol, it_ref=info["iterations"], it_ours=st["iterations"],
                   it_ref_lastbit_perturbed=it_ref[1:], it_ours_lastbit_perturbed=it_ours[1:], xerr_ref_vs_ref_perturbed=ref_self,
                   nrm_r0_ref=info["nrm_r0"], nrm_r0_ours=st["nrm_r0"], relres_ref=rr_ref, relres_ours=rr_ours, xerr=xerr,
                   loop_s_ref=dtr, loop_s_ours=dt)
        _record(row)
        assert st["converged"], name
        assert abs(info["nrm_r0"] - st["nrm_r0"]) <= 1e-12 * st["nrm_r0"], name
        _judge(name, tol, it_ref, it_ours, xerr, ref_self, rr_ours, bad)
        if name.startswith("mat900"):           # the reference's well-conditioned fixture: the literal north_star bars hold
            assert abs(st["iterations"] - info["iterations"]) <= 2 and xerr <= 1e-8, (name, row)
    assert not bad, bad


def test_shifted_vs_reference_bicgstab(cm, ref, pin):
    """reference shifted entry (pbicgstab.cu:926): valid only for n <= 524288 (mult_spec launch bug, :645)."""
    O = ref
    bad = []
    cases = []
    # the reference's own test_A0_d (example.cpp:33-106)
    ia, ja, a0 = pin["mat3_A0_ia"], pin["mat3_A0_ja"], pin["mat3_A0_a"]
    d = O.to_dense_vector(3, pin["vec3_d_a"], pin["vec3_d_ia"])
    b = O.to_dense_vector(3, pin["vec3_a"], pin["vec3_ia"])
    cases.append(("mat3_A0+d/1e-5", ia, ja, a0, d, np.ones(3), b, 1e-5))
    rng = np.random.default_rng(3)
    for nm in ("mat900", "mat10000"):
        ia, ja, a = pin[nm + "_ia"], pin[nm + "_ja"], pin[nm + "_a"]
        n = len(ia) - 1
        for tol in (1e-6, 1e-10):
            cases.append(("%s/d=0/ones/%g" % (nm, tol), ia, ja, a, np.zeros(n), np.ones(n), np.ones(n), tol))
        cases.append(("%s/d=rand/glibc_rand/1e-10" % nm, ia, ja, a, rng.uniform(0.0, 1.0, n), np.ones(n), O.glibc_rand_vector(n), 1e-10))
    for N in (32, 64):
        ia, ja, a, b = _poisson_case(O, N)
        cases.append(("poisson%d/d=0/Axtrue/1e-10" % N, ia, ja, a, np.zeros(N ** 3), np.ones(N ** 3), b, 1e-10))
    for name, ia, ja, a0, d, x0, b, tol in cases:
        base = int(ia[0])
        (xr, dtr, info), (x, dt, st), it_ref, it_ours, ref_self = _band(
            lambda bb: O.ref_gpu_bicgstab_shifted(ia, ja, a0, d, x0, bb, maxit=2000, tol=tol),
            lambda bb: cm.bicgstab_shifted(a0, ia, ja, d, x0, bb, maxit=2000, tol=tol), b)
        ia0, ja0 = ia - base, ja - base
        nrm0 = _relres(O, ia0, ja0, a0, b, x0, d=d)
        rr_ours = _relres(O, ia0, ja0, a0, b, x, d=d) / nrm0
        rr_ref = _relres(O, ia0, ja0, a0, b, xr, d=d) / nrm0
        xerr = float(np.linalg.norm(x - xr) / np.linalg.norm(xr))
        # reference: k counts from 0 and the converged iteration is printed before returning => iterations = #lines;
        # ours reports the loop counter at exit the same way (cudamat_stats.iterations)
        row = dict(case=name, mode="shifted", n=len(b), tol=tol, it_ref=info["iterations"], it_ours=st["iterations"],
                   it_ref_lastbit_perturbed=it_ref[1:], it_ours_lastbit_perturbed=it_ours[1:], xerr_ref_vs_ref_perturbed=ref_self,
                   ref_returned=info["returned"], nrm_r0_ref=info["nrm_r0"], nrm_r0_ours=st["nrm_r0"], relres_ref=rr_ref,
                   relres_ours=rr_ours, xerr=xerr, loop_s_ref=dtr, loop_s_ours=dt)
        _record(row)
        assert st["converged"] and info["returned"], name
        _judge(name, tol, it_ref, it_ours, xerr, ref_self, rr_ours, bad)
        if name.startswith("mat3") or name.startswith("mat900"):
            assert abs(st["iterations"] - info["iterations"]) <= 2 and xerr <= 1e-8, (name, row)
    assert not bad, bad


def test_reference_plain_entry_is_broken_ours_is_not(cm, ref, pin):
    """pbicgstab.cu:469-478: r0 is never set, so the reference's plain bicgstab(A,b) fails on every input; MODE_PLAIN
    implements the intended algorithm (= the shifted entry with d = 0, x0 = ones).  Documents the deviation."""
    O = ref
    ia, ja, a = pin["mat900_ia"], pin["mat900_ja"], pin["mat900_a"]
    b = np.ones(900)
    xr, _, info = O.ref_gpu_bicgstab_plain(ia, ja, a, b, maxit=50, tol=1e-6)
    assert not info["returned"]
    x, _, st = cm.bicgstab(a, ia, ja, b, maxit=2000, tol=1e-6)
    xs, _, infos = O.ref_gpu_bicgstab_shifted(ia, ja, a, np.zeros(900), np.ones(900), b, maxit=2000, tol=1e-6)
    assert st["converged"] and infos["returned"]
    assert abs(st["iterations"] - infos["iterations"]) <= 2
    assert np.linalg.norm(x - xs) <= 1e-4 * np.linalg.norm(xs)
