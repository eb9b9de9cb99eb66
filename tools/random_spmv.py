#!/usr/bin/env python3
"""dev tool: plain y = A x on the generator.cpp-style random matrix (BASELINE config 4) for the CSR variants, CUDA-event timed.
usage: python tools/random_spmv.py [--rows 50000000] [--variants 1,7] [--reps 5]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=50_000_000)
    ap.add_argument("--variants", default="1,7")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--blocks", default="", help="comma list of stream_blocks values to try with variant 7")
    ap.add_argument("--l2-fetch", default="", help="comma list of cudaLimitMaxL2FetchGranularity values to try (32,64,128)")
    args = ap.parse_args()
    import torch
    cm = ge.load_package()
    n = args.rows
    f64 = dict(dtype=torch.float64, device="cuda")
    ia = torch.empty(n + 1, dtype=torch.int32, device="cuda")
    nnz = cm.gen_random_dd_device(n, 20240, ia.data_ptr())
    ja = torch.empty(nnz, dtype=torch.int32, device="cuda")
    a = torch.empty(nnz, **f64)
    cm.gen_random_dd_device(n, 20240, ia.data_ptr(), ja.data_ptr(), a.data_ptr())
    x = torch.empty(n, **f64)
    cm.gen_xtrue_device(1234, 0, n, x.data_ptr())
    y = torch.empty(n, **f64)
    ref = None
    cases = [(int(t), None) for t in args.variants.split(",")]
    if args.l2_fetch:
        cases = [(v, int(g)) for g in args.l2_fetch.split(",") for v, _ in cases]
    if args.blocks:
        cases = [(v, None) for v, _ in cases if v != 7] + [(7, -int(k)) for k in args.blocks.split(",")]
    for v, gran in cases:
        s = cm.Solver(n)
        if gran is not None and gran < 0:
            s.set_option("stream_blocks", -gran)
            print("stream_blocks %d:" % -gran, end=" ")
            gran = None
        if gran is not None:
            s.set_option("l2_fetch", gran)
            print("L2 fetch granularity %d B:" % gran, end=" ")
        s.set_option("spmv_variant", v)
        s.set_csr_device(nnz, a.data_ptr(), ia.data_ptr(), ja.data_ptr(), keep=(ia, ja, a))
        st = s.analyze(0)
        s.spmv(x.data_ptr(), y.data_ptr(), variant=v)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            s.spmv(x.data_ptr(), y.data_ptr(), variant=v)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.reps
        if ref is None:
            ref = y.clone()
        same = bool(torch.equal(ref, y))
        print("variant %d (planned %d): %.3f ms  %.0f GB/s CSR-algorithmic  nnz=%d  bit-identical to first=%s"
              % (v, st["spmv_variant"], ms, (12 * nnz + 4 * (n + 1) + 16 * n) / ms / 1e6, nnz, same), flush=True)
        s.close()


if __name__ == "__main__":
    main()
