#!/usr/bin/env python3
"""SpMV / iteration micro-benchmark on the Poisson N^3 system (GPU only): times the SpMV kernels (with their fused
dot epilogues) inside the unpreconditioned loop with CUDA events, for each variant / stage count.
usage: python tools/spmv_bench.py [--grid 256] [--iters 60] [--configs 1:0,2:0,2:3 ...]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=256)
    ap.add_argument("--iters", type=int, default=60)
    ap.add_argument("--configs", default="1:0,2:2,2:3,2:4,2:0")
    ap.add_argument("--mode", type=int, default=0)
    ap.add_argument("--perturb", action="store_true", help="variable-coefficient stencil (values differ per row)")
    args = ap.parse_args()
    import torch
    cm = ge.load_package()
    N = args.grid
    n = N ** 3
    nnz = cm.poisson3d_nnz(N)
    f64 = dict(dtype=torch.float64, device="cuda")
    ia = torch.empty(n + 1, dtype=torch.int32, device="cuda")
    ja = torch.empty(nnz, dtype=torch.int32, device="cuda")
    a = torch.empty(nnz, **f64)
    cm.gen_poisson3d_device(N, 0, n, ia.data_ptr(), ja.data_ptr(), a.data_ptr())
    if args.perturb:        # variable coefficients: the offset dictionary still applies, the value dictionary does not
        g = torch.Generator(device="cuda"); g.manual_seed(1)
        a *= 1.0 + 0.01 * torch.rand(nnz, generator=g, **f64)
    xt = torch.empty(n, **f64)
    cm.gen_xtrue_device(1234, 0, n, xt.data_ptr())
    b = torch.empty(n, **f64)
    x = torch.zeros(n, **f64)
    bspmv = 12 * nnz + 4 * (n + 1) + 16 * n
    biter = 2 * bspmv + 120 * n
    for cfg in args.configs.split(","):
        parts = [int(v) for v in cfg.split(":")] + [0, 0, 0]
        variant, stages, tctas, cctas = parts[0], parts[1], parts[2], parts[3]
        s = cm.Solver(n)
        s.set_option("spmv_variant", variant)
        if stages:
            s.set_option("staged_stages", stages)
        if tctas:
            s.set_option("sptrsv_ctas_per_sm", tctas)
        if cctas:
            s.set_option("class_tiles_per_cta", cctas)
        s.set_csr_device(nnz, a.data_ptr(), ia.data_ptr(), ja.data_ptr())
        sa = s.analyze(args.mode)
        s.spmv(xt.data_ptr(), b.data_ptr())
        s.solve(args.mode, b.data_ptr(), x.data_ptr(), maxit=5, tol=0.0)
        # plain SpMV (no dot epilogue)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            s.spmv(xt.data_ptr(), b.data_ptr())
        e1.record()
        torch.cuda.synchronize()
        plain_ms = e0.elapsed_time(e1) / 20
        s.spmv(xt.data_ptr(), b.data_ptr())
        best = None
        for _ in range(3):
            st = s.solve(args.mode, b.data_ptr(), x.data_ptr(), maxit=args.iters, tol=0.0)
            t = st["t_loop"] * 1e3 / max(st["iterations"], 1)
            best = t if best is None else min(best, t)
        ms_it = best
        s.set_option("time_spmv", 1)
        st = s.solve(args.mode, b.data_ptr(), x.data_ptr(), maxit=args.iters, tol=0.0)
        ms_it_timed = st["t_loop"] * 1e3 / max(st["iterations"], 1)
        sp_ms = st["t_spmv"] * 1e3 / max(st["n_spmv"], 1)
        print(json.dumps({"grid": N, "variant": sa["spmv_variant"], "stages_opt": stages, "sptrsv_ctas": tctas, "class_ctas": cctas, "lib": os.path.basename(cm.LIB_PATH), "plain_spmv_ms": round(plain_ms, 4),
                          "plain_spmv_GBps": round(bspmv / plain_ms / 1e6, 1), "loop_spmv_ms": round(sp_ms, 4),
                          "loop_spmv_GBps": round(bspmv / sp_ms / 1e6, 1), "ms_per_iter": round(ms_it, 4), "ms_per_iter_with_events": round(ms_it_timed, 4),
                          "iters_per_s": round(1e3 / ms_it, 1), "iter_GBps": round(biter / ms_it / 1e6, 1),
                          "iterations": st["iterations"]}))
        s.close()


if __name__ == "__main__":
    main()
