"""GPU tests of the persistent cooperative iteration kernel (csrc/persist.cu; pytest -m gpu): the unpreconditioned loop as one
kernel per batch of iterations with grid-wide barriers — used automatically for small systems / shards, forced here.  Same
arithmetic spec, so x, the residual history and the iteration count must equal the CPU oracle's bit for bit."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _dev(torch, a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _cases(O, pin):
    out = {}
    out["mat900"] = (pin["mat900_ia"] - 1, pin["mat900_ja"] - 1, pin["mat900_a"])          # 9-point, class dictionary
    out["mat10000"] = (pin["mat10000_ia"] - 1, pin["mat10000_ja"] - 1, pin["mat10000_a"])  # 5-point, class dictionary
    out["poisson40"] = O.poisson3d(40)                                                       # 64000 rows: 32 tiles, ragged last tile
    out["random_dd"] = O.random_dd(6000, 20240)                                              # CSR row sums, rows of > 32 entries
    return out


@pytest.mark.parametrize("name", ["mat900", "mat10000", "poisson40", "random_dd"])
def test_persistent_loop_bit_identical_to_oracle(cm, O, pin, torch_cuda, name):
    torch = torch_cuda
    ia, ja, a = _cases(O, pin)[name]
    n = len(ia) - 1
    rng = np.random.default_rng(12)
    xt = O.xtrue(5, 0, n)
    b = O.spmv(ia, ja, a, xt)
    for use_d in (False, True):
        d = rng.uniform(0.0, 0.5, n) if use_d else None
        x0 = rng.standard_normal(n) if use_d else None
        xo, so = O.bicgstab_unprec(ia, ja, a, b, d=d, x0=x0, maxit=3000, tol=1e-10)
        results = []
        for persist in (1, 0):
            s = cm.Solver(n, stream=torch.cuda.current_stream().cuda_stream)
            s.set_option("persist", persist)
            s.set_csr_host(a, ia, ja)
            s.analyze(cm.MODE_SHIFTED if use_d else cm.MODE_PLAIN)
            db, dx = _dev(torch, b), torch.zeros(n, dtype=torch.float64, device="cuda")
            kw = dict(d_x0=_dev(torch, x0).data_ptr(), d_d=_dev(torch, d).data_ptr()) if use_d else {}
            keep = (_dev(torch, x0), _dev(torch, d)) if use_d else None
            if use_d:
                kw = dict(d_x0=keep[0].data_ptr(), d_d=keep[1].data_ptr())
            st = s.solve(cm.MODE_SHIFTED if use_d else cm.MODE_PLAIN, db.data_ptr(), dx.data_ptr(), maxit=3000, tol=1e-10, **kw)
            torch.cuda.synchronize()
            assert st["fused"] == (4 if persist else st["fused"] & 3), st
            assert st["converged"] == so["converged"] and st["iterations"] == so["iterations"], (name, use_d, persist, st["iterations"], so["iterations"])
            assert np.array_equal(dx.cpu().numpy(), xo), (name, use_d, persist)
            assert np.array_equal(s.history(), so["hist"]), (name, use_d, persist)
            results.append(st["t_loop"])
            s.close()
        print("persist %-10s d=%-5s %d iterations: persistent %.3f ms, per-kernel %.3f ms" % (name, use_d, so["iterations"], results[0] * 1e3, results[1] * 1e3))


def test_persistent_loop_maxit_resume_and_auto(cm, O, pin, torch_cuda):
    torch = torch_cuda
    ia, ja, a = pin["mat10000_ia"] - 1, pin["mat10000_ja"] - 1, pin["mat10000_a"]
    n = len(ia) - 1
    b = np.ones(n)
    s = cm.Solver(n, stream=torch.cuda.current_stream().cuda_stream)          # automatic: a 10^4-row system is latency bound
    s.set_csr_host(a, ia, ja)
    s.analyze(cm.MODE_PLAIN)
    db, dx = _dev(torch, b), torch.zeros(n, dtype=torch.float64, device="cuda")
    st = s.solve(cm.MODE_PLAIN, db.data_ptr(), dx.data_ptr(), maxit=21, tol=0.0)          # not a multiple of the batch of 8
    assert st["fused"] == 4 and st["iterations"] == 21 and not st["converged"] and st["breakdown"] == 3
    x21 = dx.clone()
    xo, so = O.bicgstab_unprec(ia, ja, a, b, maxit=21, tol=0.0)
    assert np.array_equal(x21.cpu().numpy(), xo) and np.array_equal(s.history(), so["hist"])
    s.solve(cm.MODE_PLAIN, db.data_ptr(), dx.data_ptr(), maxit=9, tol=0.0)
    s.set_option("resume", 1)
    st = s.solve(cm.MODE_PLAIN, db.data_ptr(), dx.data_ptr(), maxit=12, tol=0.0)
    torch.cuda.synchronize()
    assert st["iterations"] == 21 and torch.equal(dx, x21)
    s.set_option("resume", 0)
    st = s.solve(cm.MODE_PLAIN, db.data_ptr(), dx.data_ptr(), maxit=0, tol=1e-6)
    assert st["iterations"] == 0 and not st["converged"]
    s.close()
