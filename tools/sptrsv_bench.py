#!/usr/bin/env python3
"""Times the ILU0 triangular sweeps (sync-free and level-per-launch) on Poisson N^3 through the C ABI."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grids", default="64,128,256")
    ap.add_argument("--ctas", default="0")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--blocked", default="1,0", help="block-wavefront sweeps (csrc/sweepblk.cu) on / off")
    ap.add_argument("--levels", action="store_true", help="also time the level-per-launch schedule")
    args = ap.parse_args()
    import torch
    cm = ge.load_package()
    for N in [int(v) for v in args.grids.split(",")]:
        n = N ** 3
        nnz = cm.poisson3d_nnz(N)
        f64 = dict(dtype=torch.float64, device="cuda")
        ia = torch.empty(n + 1, dtype=torch.int32, device="cuda")
        ja = torch.empty(nnz, dtype=torch.int32, device="cuda")
        a = torch.empty(nnz, **f64)
        cm.gen_poisson3d_device(N, 0, n, ia.data_ptr(), ja.data_ptr(), a.data_ptr())
        rhs = torch.empty(n, **f64)
        cm.gen_xtrue_device(7, 0, n, rhs.data_ptr())
        out = torch.zeros(n, **f64)
        for blocked in [int(v) for v in args.blocked.split(",")]:
          for ctas in [int(v) for v in args.ctas.split(",")]:
            for syncfree in ([1, 0] if args.levels else [1]):
                s = cm.Solver(n)
                s.set_option("sptrsv_blocked", blocked)
                s.set_csr_device(nnz, a.data_ptr(), ia.data_ptr(), ja.data_ptr())
                sa = s.analyze(cm.MODE_ILU0)
                s.set_option("sptrsv_syncfree", syncfree)
                if ctas:
                    s.set_option("sptrsv_ctas_per_sm", ctas)
                res = {}
                for upper in (0, 1):
                    s.sptrsv(upper, rhs.data_ptr(), out.data_ptr())
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(args.reps):
                        s.sptrsv(upper, rhs.data_ptr(), out.data_ptr())
                    e1.record()
                    torch.cuda.synchronize()
                    res["U" if upper else "L"] = e0.elapsed_time(e1) / args.reps
                # chained: U sweep reading the vector the L sweep has just scatter-written (as in the solver loop)
                out2 = torch.zeros(n, **f64)
                s.sptrsv(0, rhs.data_ptr(), out.data_ptr()); s.sptrsv(1, out.data_ptr(), out2.data_ptr())
                torch.cuda.synchronize()
                e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
                tl_, tu_ = 0.0, 0.0
                for _ in range(args.reps):
                    e0.record(); s.sptrsv(0, rhs.data_ptr(), out.data_ptr()); e1.record()
                    s.sptrsv(1, out.data_ptr(), out2.data_ptr()); e2.record()
                    torch.cuda.synchronize()
                    tl_ += e0.elapsed_time(e1); tu_ += e1.elapsed_time(e2)
                res["chainL"], res["chainU"] = tl_ / args.reps, tu_ / args.reps
                print(json.dumps({"grid": N, "blocked": blocked, "chain_L_ms": round(res["chainL"], 4), "chain_U_ms": round(res["chainU"], 4), "levels": sa["levels_l"], "syncfree": syncfree, "ctas_per_sm": ctas,
                                  "L_ms": round(res["L"], 4), "U_ms": round(res["U"], 4),
                                  "us_per_level": round(res["L"] * 1e3 / sa["levels_l"], 3),
                                  "ns_per_row": round(res["L"] * 1e6 / n, 3), "t_analysis_s": round(sa["t_analysis"], 3)}))
                s.close()


if __name__ == "__main__":
    main()
