#!/usr/bin/env python3
"""bench.py — BiCGSTAB iterations/s + SpMV HBM GB/s on the 3-D 7-point Poisson 256^3 system (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--grid 256]

A "step" is ONE BiCGSTAB iteration (2 SpMV + the fused vector updates / reductions) of the unpreconditioned
loop on the Poisson N^3 system, inputs resident in HBM.  The K timed steps run as ceil(K/chunk) restarts of
`chunk` iterations from x0 = ones (tol = 0 so nothing stops early; the one-time residual set-up of each
restart, 1 SpMV per chunk, is inside the timed region).  Matrix + vectors (>= 2.4 GB at 256^3) are far
larger than the 126 MB L2, so no explicit L2 flush is needed.

N > 1 (torchrun, one rank per GPU): the SAME global system row-sharded over the ranks (strong scaling).

JSON line keys: see the task contract; extras: "converge" (full solve to 1e-10), "roofline", "cpu_baseline",
"clocks", "e2e" (host-pointer C-ABI solve incl. H2D/D2H), "ilu0" and "mat10000" (other BASELINE configs).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

METRIC = "BiCGSTAB iters/s + SpMV HBM GB/s (% peak), 3D Poisson 256^3 at 1/2/4/8 B200"
UNIT = "iterations/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def bytes_spmv(n, nnz):
    return 12 * nnz + 4 * (n + 1) + 16 * n


def bytes_iter(n, nnz):
    return 2 * bytes_spmv(n, nnz) + 120 * n


class ClockSampler:
    """samples SM clock / throttle reasons of one GPU with NVML while the timed region runs"""

    def __init__(self, index=0, period=0.05):
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:      # noqa: BLE001
            self.nv, self.err = None, str(e)

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake": 0x80, "sync_boost": 0x10, "applications_clocks": 0x2}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:       # noqa: BLE001
                pass
            time.sleep(self.period)

    def __enter__(self):
        if self.nv:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t:
            self._t.join()

    def summary(self):
        if not self.nv or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the reference's own CPU implementation (bicstab_omp BiCG) on host cores
# ---------------------------------------------------------------------------------------------------
def host_cores():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        return max(1, os.cpu_count() or 1)


def cpu_reference_run(N, budget_s=25.0, iters=None):
    """Times the reference's bicstab_omp BiCG() (oracle/_ref, compiled from the reference sources) on the
    Poisson N^3 system, x0 = ones. BiCG() transposes the matrix inside the call, so the rate is taken from
    the difference of two calls with different iteration caps.  Falls back to the oracle port (1 core)."""
    O = ge.load_oracle()
    ia, ja, a = O.poisson3d(N)
    n = N ** 3
    xt = O.xtrue(1234, 0, n)
    b = O.spmv(ia, ja, a, xt)
    if O.ref_available("bicg"):
        cores = O.ref_omp_threads(host_cores())
        O.ref_bicg(ia, ja, a, b, maxit=1)                      # untimed: library load, OpenMP pool, first-touch pages
        t1 = None
        for _ in range(2):                                     # T(maxit=1) = set-up + transpose + 1 iteration (best of 2)
            t0 = time.time(); O.ref_bicg(ia, ja, a, b, maxit=1); dt = time.time() - t0
            t1 = dt if t1 is None else min(t1, dt)
        t0 = time.time(); O.ref_bicg(ia, ja, a, b, maxit=5); t5 = time.time() - t0
        per_it = max((t5 - t1) / 4.0, 1e-6)
        # at least 10 timed iterations whatever --steps says: a shorter difference of two calls is noise
        k2 = max(11, iters) if iters else int(max(11, min(200, (budget_s - t1) / per_it)))
        t0 = time.time(); _, it_done = O.ref_bicg(ia, ja, a, b, maxit=k2); tk = time.time() - t0
        rate = (it_done - 1) / max(tk - t1, (it_done - 1) * per_it * 0.5, 1e-9)
        return {"value": rate, "unit": UNIT, "cores": cores, "kind": "reference",
                "sample": "reference bicstab_omp BiCG() (BiCG, 2 SpMV/iteration incl. A^T) on Poisson %d^3, x0=ones: "
                          "(%d-1) iterations / (T(maxit=%d)-T(maxit=1)) = %.2f s; OMP threads=%d"
                          % (N, it_done, k2, tk - t1, cores)}, (it_done - 1), (tk - t1)
    # oracle port of the BiCGSTAB loop, scalar, 1 core
    k = iters if iters else 3
    t0 = time.time(); O.bicgstab_unprec(ia, ja, a, b, maxit=k, tol=0.0); tk = time.time() - t0
    return {"value": k / tk, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "oracle port of the unpreconditioned BiCGSTAB loop on Poisson %d^3, %d iterations, 1 core" % (N, k)}, k, tk


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 to its children: the reference arm must use the box's host cores whatever the
    # launcher (set before libgomp is loaded by the shim; also forced through omp_set_num_threads below)
    os.environ["OMP_NUM_THREADS"] = str(host_cores())
    cb, its, secs = cpu_reference_run(args.grid, budget_s=60.0, iters=None if args.steps <= 0 else min(args.steps, 200))
    n = args.grid ** 3
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / cb["value"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "poisson3d_%d" % args.grid, "n": n,
                       "mode": "reference CPU path: bicstab_omp BiCG (NOT BiCGSTAB: 2 SpMV per iteration incl. A^T, bicstab.cpp:93-196), OpenMP threads = %d" % cb["cores"],
                       "timed_iterations": its},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
VNAME = {1: "k_spmv_rowlane (CSR)", 2: "k_spmv_staged (CSR, TMA ring)", 3: "k_spmv_class<values from CSR>",
         4: "k_spmv_class<values from dictionary>", 5: "k_spmv_tiled<class dictionary, x windows staged by TMA>",
         6: "k_spmv_march<plane-marching ring, class dictionary>"}


def kernel_bytes(variant, fused, nloc, nnz_loc):
    """bytes each timed kernel of ONE iteration must move in the storage format it really reads (per rank), by timing slot
    (cudamat_stats.t_kernel): 0 = SpMV 1 (+ folded p update), 1 = SpMV 2 (+ folded s update), 2 = x/r update + 2 dots,
    3 = separate p and s updates (average of the two launches).  Dictionary formats read 1 B per row instead of 12 B per entry."""
    if variant in (4, 5, 6):
        mat = nloc                                  # class id / presence mask, 1 B per row
    elif variant == 3:
        mat = 8 * nnz_loc + nloc + 4 * (nloc // 32 + 1)
    else:
        mat = 12 * nnz_loc + 4 * (nloc + 1)
    out = {2: ("k_update_xr: x, r update + rhat.r, r.r", 56 * nloc)}
    out[0] = (("MAKE_P: p'=r+beta(p-omega v); v'=A p'; rhat.v'", mat + 8 * nloc * 6) if fused & 1          # r, p, v, rhat in; p', v' out
              else ("SpMV 1 + rhat.v", mat + 8 * nloc * 3))                                                 # x, rhat in; y out
    out[1] = (("MAKE_S: s=r-alpha v; t=A s; t.s, t.t", mat + 8 * nloc * 4) if fused & 2                     # r, v in; s, t out
              else ("SpMV 2 + t.s, t.t", mat + 8 * nloc * 2))
    if fused != 3:
        sep = [b for b, f in ((32 * nloc, 1), (24 * nloc, 2)) if not fused & f]
        out[3] = ("k_update_p / k_update_s (separate launches: %d per iteration)" % len(sep), sum(sep) // len(sep))
    return out


def run_ours(args):
    import hashlib
    import torch
    import torch.distributed as dist
    cm = ge.load_package()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available() or cm.device_count() <= 0:
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    N = args.grid
    n = N ** 3
    row0, row1 = cm.partition_rows(n, world, rank)
    nloc = row1 - row0
    nnz_loc = cm.poisson3d_nnz(N, row0, row1)
    nnz = cm.poisson3d_nnz(N)
    stream = torch.cuda.current_stream().cuda_stream
    f64 = dict(dtype=torch.float64, device="cuda")
    ia = torch.empty(nloc + 1, dtype=torch.int32, device="cuda")
    ja = torch.empty(nnz_loc, dtype=torch.int32, device="cuda")
    a = torch.empty(nnz_loc, **f64)
    cm.gen_poisson3d_device(N, row0, row1, ia.data_ptr(), ja.data_ptr(), a.data_ptr(), stream)

    def new_solver(variant=0):
        sol = cm.Solver(n, row0, row1, stream=stream)
        sol.set_csr_device(nnz_loc, a.data_ptr(), ia.data_ptr(), ja.data_ptr(), keep=(ia, ja, a))
        if world > 1:
            idbuf = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if rank == 0:
                idbuf = torch.tensor(list(cm.Comm.unique_id()), dtype=torch.uint8, device="cuda")
            dist.broadcast(idbuf, 0)
            cm.Comm.init(sol, bytes(idbuf.cpu().tolist()), rank, world)
        if variant:
            sol.set_option("spmv_variant", variant)
        if args.fuse >= 0:
            sol.set_option("fuse", args.fuse)
        if args.persist >= -1:
            sol.set_option("persist", args.persist)
        return sol, sol.analyze(cm.MODE_PLAIN)

    s, sa = new_solver(args.variant)
    xt = torch.empty(nloc, **f64)
    cm.gen_xtrue_device(1234, row0, nloc, xt.data_ptr(), stream)
    b = torch.empty(nloc, **f64)
    s.spmv(xt.data_ptr(), b.data_ptr())          # b = A x_true (distributed SpMV when sharded)
    x = torch.zeros(nloc, **f64)
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    clk = ClockSampler(local_rank, period=0.004)
    clk.__enter__()                               # samples from here to the end of the timed region (the GPU is under load throughout)
    # ---- full solve to 1e-10: the convergence claim of the metric (also part of the warm-up) ----
    barrier()
    t0 = time.time()
    stc = s.solve(cm.MODE_PLAIN, b.data_ptr(), x.data_ptr(), maxit=5 if args.no_converge else 5000, tol=1e-10)
    torch.cuda.synchronize()
    t_conv = time.time() - t0
    err = torch.stack([torch.sum((x - xt) ** 2), torch.sum(xt ** 2)])
    if world > 1:
        dist.all_reduce(err)
    relerr = float(torch.sqrt(err[0] / err[1]))
    converge = {"tol": 1e-10, "iterations": stc["iterations"], "converged": bool(stc["converged"]),
                "relres": stc["nrm_r"] / stc["nrm_r0"], "rel_err_vs_xtrue": relerr, "seconds": t_conv,
                "iters_per_s": stc["iterations"] / stc["t_loop"]}
    # bit-identity evidence: sha256 of the full solution (gathered in rank order) against the committed digest of the CPU
    # oracle's own run (tests/golden/poisson<N>_oracle_digest.json) - the same digest whatever the number of GPUs
    if not args.no_converge:
        try:
            if world > 1:
                cnt = torch.tensor([nloc], dtype=torch.int64, device="cuda")
                cnts = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
                dist.all_gather(cnts, cnt)
                mx = int(max(int(c[0]) for c in cnts))
                pad = torch.zeros(mx, **f64); pad[:nloc] = x
                parts = [torch.empty(mx, **f64) for _ in range(world)] if rank == 0 else None
                dist.gather(pad, parts, dst=0)
                xfull = torch.cat([parts[r][:int(cnts[r][0])] for r in range(world)]).cpu().numpy() if rank == 0 else None
            else:
                xfull = x.cpu().numpy()
            if rank == 0:
                converge["x_sha256"] = hashlib.sha256(np.ascontiguousarray(xfull).tobytes()).hexdigest()
                dp = os.path.join(ROOT, "tests", "golden", "poisson%d_oracle_digest.json" % N)
                if os.path.exists(dp):
                    dg = json.load(open(dp))
                    converge["oracle_digest"] = {"file": "tests/golden/poisson%d_oracle_digest.json" % N, "iterations": dg["iterations"],
                                                 "x_sha256": dg["x_sha256"]}
                    converge["bit_identical_to_cpu_oracle"] = bool(dg["x_sha256"] == converge["x_sha256"] and dg["iterations"] == stc["iterations"])
                del xfull
        except Exception as e:      # noqa: BLE001
            converge["x_sha256_error"] = str(e)

    chunk = max(8, min(args.chunk, converge["iterations"] // 2 if converge["converged"] else args.chunk))
    peak, peak_src = peaks()
    b_spmv = bytes_spmv(n, nnz)           # whole-problem algorithmic bytes of one CSR SpMV
    b_it = bytes_iter(n, nnz)

    def timed_region(sol, K, W):
        """EXACTLY K iterations between the two events (CUDA events on the launching stream, max over ranks).  The warm-up
        solve (residual set-up + W iterations) runs before the first event; the timed window CONTINUES that solve (option
        "resume") so no set-up work sits inside a short window, and restarts from x0 = ones every `chunk` iterations
        (tol = 0: nothing stops early; a restart's 1-SpMV residual set-up is inside the window, 1 per `chunk` iterations)."""
        W = max(W, 3)
        first = max(1, min(K, chunk - min(W, chunk - 1)))
        plan = [first] + [chunk] * ((K - first) // chunk) + ([(K - first) % chunk] if (K - first) % chunk else [])
        sol.set_option("resume", 0)
        launches0 = sol.solve(cm.MODE_PLAIN, b.data_ptr(), x.data_ptr(), maxit=W, tol=0.0)["kernel_launches"]
        # the main kernels of every 2nd (8th) iteration are event-timed.  Sharded runs always use every 8th: an event-timed kernel
        # leaves the PDL chain, its halo push comes late and the neighbour waits for it (2 GPUs, K = 20: 0.329 ms per step with
        # every 2nd iteration instrumented against 0.293-0.301 with every 8th)
        sol.set_option("time_spmv", 2 if (K < 64 and world == 1) else 8)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tk, nk, done, launches, prev = [0.0] * 4, [0] * 4, 0, 0, W
        barrier()
        e0.record()
        for i, m in enumerate(plan):
            sol.set_option("resume", 1 if i == 0 else 0)
            st = sol.solve(cm.MODE_PLAIN, b.data_ptr(), x.data_ptr(), maxit=m, tol=0.0)
            done += st["iterations"] - (prev if i == 0 else 0)
            for q in range(4):
                tk[q] += st["t_kernel"][q]; nk[q] += st["n_kernel"][q]
            launches = st["kernel_launches"] - launches0     # the handle counts every launch since its creation
            fused = int(st["fused"])
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if done != K:
            raise SystemExit("timed region ran %d iterations instead of %d (break-down?)" % (done, K))
        sol.set_option("time_spmv", 0)
        sol.set_option("resume", 0)
        kms = [(tk[q] * 1e3 / nk[q]) if nk[q] else None for q in range(4)]
        return float(ms[0]), kms, nk, launches, fused

    def roofline(variant, fused, ms, K, kms, nk):
        """`frac` = bytes the dominant timed kernel must move in its own storage format / its average launch duration / peak."""
        kb = kernel_bytes(variant, fused, nloc, nnz_loc)
        per = ms / K
        rows = []
        for q, (nm, by) in kb.items():
            if kms[q]:
                launches_per_it = (2 - bin(fused & 3).count("1")) if q == 3 else 1
                rows.append({"slot": q, "kernel": nm, "bytes_per_launch": by, "avg_launch_ms": kms[q], "launches_timed": nk[q],
                             "GBps": by / (kms[q] * 1e-3) / 1e9, "frac": by / (kms[q] * 1e-3) / 1e9 / peak,
                             "share_of_step": launches_per_it * kms[q] / per})
        if not rows:
            # persistent cooperative iteration kernel (small shards): the whole iteration is ONE kernel, timed as ms / K
            by = sum(v[1] * (2 if q == 3 else 1) for q, v in kernel_bytes(variant, 0, nloc, nnz_loc).items())
            rows = [{"slot": -1, "kernel": "k_bicgstab_persist: whole iteration (p, SpMV 1, s, SpMV 2, x/r update, 3 reductions) in one cooperative kernel",
                     "bytes_per_launch": by, "avg_launch_ms": per, "launches_timed": K, "GBps": by / (per * 1e-3) / 1e9,
                     "frac": by / (per * 1e-3) / 1e9 / peak, "share_of_step": 1.0}]
        dom = max(rows, key=lambda r: r["share_of_step"])
        its = K / (ms * 1e-3)
        it_bytes = sum(r["bytes_per_launch"] * ((2 - bin(fused & 3).count("1")) if r["slot"] == 3 else 1) for r in rows) + 3 * 16 * (nloc // 32)
        out = {"bound": "hbm", "kernel": dom["kernel"] + " [" + VNAME.get(variant, "?") + "]",
               "achieved": dom["GBps"], "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": dom["frac"],
               "frac_of_8TBps_datasheet": dom["GBps"] / 8000.0, "traffic": None,
               "bytes_per_launch": dom["bytes_per_launch"], "avg_launch_ms": dom["avg_launch_ms"], "launches_timed": dom["launches_timed"],
               "bytes_definition": "bytes the kernel must move in the storage format it reads (dictionary variants: 1 B/row instead of "
                                   "12 B/entry), every vector read / written once; SURVEY.md 8d CSR-algorithmic figures are under iteration.csr_*",
               "kernels": rows,
               "iteration": {"format_bytes": it_bytes, "format_GBps": it_bytes * its / 1e9, "format_frac": it_bytes * its / 1e9 / peak,
                             "format_frac_of_8TBps": it_bytes * its / 1e9 / 8000.0,
                             "csr_algorithmic_bytes": b_it // world, "csr_equivalent_GBps": b_it / world * its / 1e9,
                             "csr_equivalent_frac": b_it / world * its / 1e9 / peak}}
        tp = os.path.join(ROOT, "profiles", "r2_dram_traffic.json")
        if os.path.exists(tp) and N == 256 and world == 1:
            try:
                tr = json.load(open(tp))
                key = ("fuse%d" % fused) if fused else ("variant%d" % variant)
                ent = tr.get(key)
                if ent:
                    out["traffic"] = ent.get("dominant_kernel_dram_bytes_per_launch")
                    if ent.get("iteration_dram_bytes"):
                        out["iteration"]["dram_bytes_ncu"] = ent["iteration_dram_bytes"]
                        out["iteration"]["dram_GBps"] = ent["iteration_dram_bytes"] * its / 1e9
                        out["iteration"]["dram_frac"] = ent["iteration_dram_bytes"] * its / 1e9 / peak
                        out["iteration"]["dram_frac_of_8TBps"] = ent["iteration_dram_bytes"] * its / 1e9 / 8000.0
                        out["iteration"]["dram_source"] = ent.get("source")
            except Exception:       # noqa: BLE001
                pass
        return out

    # ---- timed region: the planned (AUTO) SpMV variant ----
    K = args.steps
    ms, kms, nk, last_launches, fused = timed_region(s, K, args.warmup)
    clk.__exit__()
    clocks = clk.summary()
    its_per_s = K / (ms * 1e-3)
    roof = roofline(sa["spmv_variant"], fused, ms, K, kms, nk)

    line = {"metric": METRIC, "value": its_per_s, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": args.warmup,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "poisson3d_%d" % N, "n": n, "nnz": nnz, "mode": "unpreconditioned BiCGSTAB (pbicgstab.h:113)",
                       "x0": "ones", "b": "A*x_true, x_true=hash(1234,i) in (-1,1)", "chunk": chunk,
                       "timed_window": "continues the warm-up solve (no set-up inside), restart every `chunk` iterations",
                       "l2": "inputs (%.1f GB CSR + vectors) larger than the 126 MB L2, no flush" % ((12 * nnz + 4 * n) / 1e9),
                       "spmv_variant": sa["spmv_variant"], "spmv_format": VNAME.get(sa["spmv_variant"]),
                       "fused_updates": fused,
                       "sharding": "row slabs, %d rank(s)" % world,
                       "comm": ("none" if world == 1 else ("peer memory over NVLink (CUDA IPC): fused halo pushes + in-kernel gather of the partial sums, no NCCL call per iteration"
                                                            if cm.Comm.p2p_enabled(s) else "NCCL send/recv halo + allreduce of the partial sums"))},
            "gpu_launches": int(last_launches),
            "converge": converge, "roofline": roof, "clocks": clocks}

    # ---- the same steps with the plain CSR kernel (variant ROWLANE): the apples-to-apples roofline of SURVEY.md 8d ----
    if sa["spmv_variant"] != 1 and not args.no_csr and roof is not None:
        s1, _ = new_solver(1)
        K1 = min(K, 400)
        ms1, kms1, nk1, _, _ = timed_region(s1, K1, args.warmup)
        r1 = roofline(1, 0, ms1, K1, kms1, nk1)
        roof["csr_kernel"] = {"value": K1 / (ms1 * 1e-3), "unit": UNIT, "steps": K1, "ms_per_step": ms1 / K1,
                              "note": "same timed window with the plain CSR kernel (12 B/entry streamed): SURVEY.md 8d's B_spmv / B_iter apply literally",
                              "spmv": [k for k in (r1["kernels"] if r1 else []) if k["slot"] in (0, 1)],
                              "iteration_GBps": r1["iteration"]["csr_equivalent_GBps"] if r1 else None,
                              "iteration_frac": r1["iteration"]["csr_equivalent_frac"] if r1 else None,
                              "iteration_frac_of_8TBps": (r1["iteration"]["csr_equivalent_GBps"] / 8000.0) if r1 else None}
        s1.close()

    if rank == 0 and world == 1 and not args.no_extras:
        line.update(single_gpu_extras(args, cm, torch, s, N, n, nnz, ia, ja, a, b, x, xt, converge))
    elif world > 1 and not args.no_extras:
        line["e2e"] = multi_gpu_e2e(cm, torch, dist, N, n, row0, row1, nnz_loc, ia, ja, a, b, rank, world, stream, barrier)
    s.close()
    # ---- BASELINE config 5: the 512^3 system (134 M rows) on the same ranks, short timed region ----
    del ia, ja, a
    torch.cuda.empty_cache()
    if not args.no_512 and N != 512:
        try:
            big = big_grid_run(cm, torch, dist, 512, world, rank, stream, barrier, steps=100)
        except Exception as e:      # noqa: BLE001
            big = {"error": str(e)}
        # inside `config` so that it survives into the driver's parsed record (SCALE: one line per N from the same code)
        line["config"]["poisson512"] = big

    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def big_grid_run(cm, torch, dist, N, world, rank, stream, barrier, steps=200):
    """iterations/s of the unpreconditioned loop on Poisson N^3 row-sharded over the ranks of this run (config 5 of
    BASELINE.json: 512^3 on 1/2/4/8 GPUs; it fits one B200).  CUDA events, max over ranks."""
    n = N ** 3
    row0, row1 = cm.partition_rows(n, world, rank)
    nloc = row1 - row0
    nnz_loc = cm.poisson3d_nnz(N, row0, row1)
    f64 = dict(dtype=torch.float64, device="cuda")
    ia = torch.empty(nloc + 1, dtype=torch.int32, device="cuda")
    ja = torch.empty(nnz_loc, dtype=torch.int32, device="cuda")
    a = torch.empty(nnz_loc, **f64)
    cm.gen_poisson3d_device(N, row0, row1, ia.data_ptr(), ja.data_ptr(), a.data_ptr(), stream)
    s = cm.Solver(n, row0, row1, stream=stream)
    s.set_csr_device(nnz_loc, a.data_ptr(), ia.data_ptr(), ja.data_ptr(), keep=(ia, ja, a))
    if world > 1:
        idbuf = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idbuf = torch.tensor(list(cm.Comm.unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(idbuf, 0)
        cm.Comm.init(s, bytes(idbuf.cpu().tolist()), rank, world)
    sa = s.analyze(cm.MODE_PLAIN)
    xt = torch.empty(nloc, **f64)
    cm.gen_xtrue_device(1234, row0, nloc, xt.data_ptr(), stream)
    b = torch.empty(nloc, **f64)
    s.spmv(xt.data_ptr(), b.data_ptr())
    x = torch.zeros(nloc, **f64)
    s.solve(cm.MODE_PLAIN, b.data_ptr(), x.data_ptr(), maxit=5, tol=0.0)
    s.set_option("resume", 1)                      # the timed window continues the warm-up solve: no set-up inside
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    st = s.solve(cm.MODE_PLAIN, b.data_ptr(), x.data_ptr(), maxit=steps, tol=0.0)
    e1.record()
    barrier()
    st["iterations"] -= 5
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms[0])
    nnz = cm.poisson3d_nnz(N)
    s.close()
    del ia, ja, a, xt, b, x
    torch.cuda.empty_cache()
    return {"workload": "poisson3d_%d" % N, "n": n, "nnz": nnz, "steps": st["iterations"], "ms_per_step": ms / max(st["iterations"], 1),
            "iters_per_s": st["iterations"] / (ms * 1e-3), "spmv_variant": sa["spmv_variant"],
            "fused_updates": int(st.get("fused", 0)), "n_gpus": world,
            "iteration_csr_GBps_per_gpu": bytes_iter(n, nnz) / world * st["iterations"] / (ms * 1e-3) / 1e9,
            "note": "strong scaling of the SAME 512^3 system over the ranks of this run (BASELINE config 5): parallel efficiency = "
                    "iters_per_s(N) / (N * iters_per_s(1)) across the driver's N = 1, 2, 4, 8 lines"}


def random_dd_run(cm, torch, n):
    """n rows, row lengths from a 90/9/1 % mixture (mean 8.5 off-diagonals, up to ~190), columns uniform
    over [0, n) (x = 400 MB does not fit L2: sector over-fetch on the gathers), values U(-10,10), dominant diagonal.
    Reports the SpMV rate (plain + inside the loop) and a full solve of A x = A x_true to 1e-10."""
    f64 = dict(dtype=torch.float64, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    ia = torch.empty(n + 1, dtype=torch.int32, device="cuda")
    nnz = cm.gen_random_dd_device(n, 20240, ia.data_ptr(), stream=stream)
    ja = torch.empty(nnz, dtype=torch.int32, device="cuda")
    a = torch.empty(nnz, **f64)
    cm.gen_random_dd_device(n, 20240, ia.data_ptr(), ja.data_ptr(), a.data_ptr(), stream=stream)
    s = cm.Solver(n, stream=stream)
    s.set_csr_device(nnz, a.data_ptr(), ia.data_ptr(), ja.data_ptr(), keep=(ia, ja, a))
    sa = s.analyze(cm.MODE_PLAIN)
    xt = torch.empty(n, **f64)
    cm.gen_xtrue_device(1234, 0, n, xt.data_ptr(), stream)
    b = torch.empty(n, **f64)
    s.spmv(xt.data_ptr(), b.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        s.spmv(xt.data_ptr(), b.data_ptr())
    e1.record()
    torch.cuda.synchronize()
    spmv_ms = e0.elapsed_time(e1) / 10
    x = torch.zeros(n, **f64)
    st = s.solve(cm.MODE_PLAIN, b.data_ptr(), x.data_ptr(), maxit=2000, tol=1e-10)
    relerr = float(torch.linalg.norm(x - xt) / torch.linalg.norm(xt))
    bs = bytes_spmv(n, nnz)
    peak, _ = peaks()
    out = {"n": n, "nnz": nnz, "max_row_len_class": "irregular (mixture 90/9/1 %)", "spmv_variant": sa["spmv_variant"],
           "spmv_ms": spmv_ms, "spmv_GBps": bs / spmv_ms / 1e6, "spmv_frac_of_measured_peak": bs / spmv_ms / 1e6 / peak,
           "iterations": st["iterations"], "converged": bool(st["converged"]), "t_loop_s": st["t_loop"],
           "iters_per_s": st["iterations"] / max(st["t_loop"], 1e-9), "relres": st["nrm_r"] / st["nrm_r0"], "rel_err_vs_xtrue": relerr,
           "iteration_GBps": bytes_iter(n, nnz) * st["iterations"] / max(st["t_loop"], 1e-9) / 1e9,
           "note": "x gathers are random over 400 MB: 32-byte sectors fetched per 8 useful bytes, the algorithmic roofline is not reachable (SURVEY.md H5)"}
    s.close()
    del ia, ja, a, xt, b, x
    torch.cuda.empty_cache()
    return out


def multi_gpu_e2e(cm, torch, dist, N, n, row0, row1, nnz_loc, ia, ja, a, b, rank, world, stream, barrier):
    """e2e at N > 1: every rank holds its row shard (CSR with global columns, b) in pinned HOST memory and goes
    through the handle API of the C ABI: cudamat_create, cudamat_set_csr_host (H2D), cudamat_comm_init (halo plan; the
    process-wide NCCL communicator already exists), cudamat_analyze, H2D of b, cudamat_solve_device to 1e-10, D2H of x.
    Wall clock between barriers, max over ranks."""
    nloc = row1 - row0
    try:
        h_ia = torch.empty(nloc + 1, dtype=torch.int32, pin_memory=True); h_ia.copy_(ia)
        h_ja = torch.empty(nnz_loc, dtype=torch.int32, pin_memory=True); h_ja.copy_(ja)
        h_a = torch.empty(nnz_loc, dtype=torch.float64, pin_memory=True); h_a.copy_(a)
        h_b = torch.empty(nloc, dtype=torch.float64, pin_memory=True); h_b.copy_(b)
        h_x = torch.empty(nloc, dtype=torch.float64, pin_memory=True)
        idbuf = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idbuf = torch.tensor(list(cm.Comm.unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(idbuf, 0)
        idb = bytes(idbuf.cpu().tolist())
        barrier()
        t0 = time.time()
        s2 = cm.Solver(n, row0, row1, stream=stream)
        s2.set_csr_host(h_a.numpy(), h_ia.numpy(), h_ja.numpy())
        t1 = time.time()
        cm.Comm.init(s2, idb, rank, world)
        t2 = time.time()
        s2.analyze(cm.MODE_PLAIN)
        t3 = time.time()
        d_b = torch.empty(nloc, dtype=torch.float64, device="cuda"); d_b.copy_(h_b, non_blocking=True)
        d_x = torch.empty(nloc, dtype=torch.float64, device="cuda")
        st = s2.solve(cm.MODE_PLAIN, d_b.data_ptr(), d_x.data_ptr(), maxit=5000, tol=1e-10)
        t4 = time.time()
        h_x.copy_(d_x)
        barrier()
        phases = {"create_upload": t1 - t0, "comm_init": t2 - t1, "analyze": t3 - t2, "solve": t4 - t3, "d2h_barrier": time.time() - t4}
        wall = torch.tensor([time.time() - t0], dtype=torch.float64, device="cuda")
        dist.all_reduce(wall, op=dist.ReduceOp.MAX)
        wall = float(wall[0])
        s2.close()
        it = st["iterations"]
        return {"value": it / wall, "unit": UNIT, "h2d_bytes_per_step": (12 * nnz_loc + 4 * (nloc + 1) + 8 * nloc) * world / max(it, 1),
                "d2h_bytes_per_step": 8 * n / max(it, 1), "iterations": it, "wall_s": wall, "t_loop_s": st["t_loop"],
                "converged": bool(st["converged"]), "phases_s_rank0": phases,
                "call": "per rank: cudamat_set_csr_host + cudamat_comm_init + cudamat_analyze + cudamat_solve_device(tol=1e-10) on pinned host shards"}
    except Exception as e:      # noqa: BLE001
        return {"value": None, "error": str(e)}


def single_gpu_extras(args, cm, torch, s, N, n, nnz, ia, ja, a, b, x, xt, converge):
    out = {}
    # ---- e2e: the reference-facing host-pointer C-ABI call (cudamat_bicgstab_host), HOST buffers, H2D of CSR + b and
    #      D2H of x inside the timed region; full solve to 1e-10.  Measured with pinned host arrays (the contract's e2e) and
    #      with pageable ones (what the reference's callers hand over: malloc'ed arrays, example.cpp:96-104,252), each as
    #      first call of the process (grows the library's device memory pool) and steady state ----
    import ctypes as C

    def host_solve(h_a, h_ia, h_ja, h_b, h_x, mode=None):
        st = cm.Stats(); dt = C.c_double(0.0)
        t0 = time.time()
        rc = cm.lib.cudamat_bicgstab_host(cm.MODE_PLAIN if mode is None else mode, n, nnz, C.cast(h_a.data_ptr(), cm.c_dp), C.cast(h_ia.data_ptr(), cm.c_ip),
                                          C.cast(h_ja.data_ptr(), cm.c_ip), None, None, C.cast(h_b.data_ptr(), cm.c_dp),
                                          5000, 1e-10, 0, C.cast(h_x.data_ptr(), cm.c_dp), C.byref(dt), C.byref(st))
        wall = time.time() - t0
        cm._check(rc)
        return wall, st

    try:
        res = {}
        for kind in ("pageable", "pinned"):
            pin = kind == "pinned"
            h_ia = torch.empty(n + 1, dtype=torch.int32, pin_memory=pin); h_ia.copy_(ia)
            h_ja = torch.empty(nnz, dtype=torch.int32, pin_memory=pin); h_ja.copy_(ja)
            h_a = torch.empty(nnz, dtype=torch.float64, pin_memory=pin); h_a.copy_(a)
            h_b = torch.empty(n, dtype=torch.float64, pin_memory=pin); h_b.copy_(b)
            h_x = torch.empty(n, dtype=torch.float64, pin_memory=pin)
            torch.cuda.synchronize()
            walls = []
            for _ in range(2):
                wall, st = host_solve(h_a, h_ia, h_ja, h_b, h_x)
                walls.append(wall)
            res[kind] = {"value": st.iterations / walls[-1], "wall_s": walls[-1], "value_first_call": st.iterations / walls[0],
                         "wall_s_first_call": walls[0], "t_h2d_s": st.t_h2d, "t_analysis_s": st.t_analysis, "t_loop_s": st.t_loop,
                         "t_d2h_s": st.t_d2h, "iterations": st.iterations, "converged": bool(st.converged)}
            del h_ia, h_ja, h_a, h_b, h_x
        h2d = 12 * nnz + 4 * (n + 1) + 8 * n
        its = res["pinned"]["iterations"]
        out["e2e"] = {"value": res["pinned"]["value"], "unit": UNIT, "h2d_bytes_per_step": h2d / max(its, 1),
                      "d2h_bytes_per_step": 8 * n / max(its, 1), "iterations": its, "wall_s": res["pinned"]["wall_s"],
                      "pinned": res["pinned"], "pageable": res["pageable"],
                      "call": "cudamat_bicgstab_host(MODE_PLAIN, tol=1e-10) on host CSR/b/x: H2D 1.54 GB + analysis + %d iterations + D2H; "
                              "`value` = pinned host arrays, steady state (second call); `pageable` = malloc'ed arrays as the reference's "
                              "callers pass them; `*_first_call` = first call of the process (the very first one also grows the device memory pool)" % its}
    except Exception as e:      # noqa: BLE001
        out["e2e"] = {"value": None, "error": str(e)}

    # ---- same-box GPU bar: the REFERENCE's own GPU code (pbicgstab.cu compiled unmodified against the test-only legacy-cuSPARSE
    #      shim, oracle/_ref) on the same 256^3 system in its only live mode (bicgstab_lu_precond, example.cpp:352), next to ours ----
    if not args.no_ilu0 and not args.no_refgpu:
        try:
            O = ge.load_oracle()
            if O.ref_available("pbicgstab"):
                h = [t.cpu().numpy() for t in (ia, ja, a, b)]
                t0 = time.time()
                xr, dtr, info = O.ref_gpu_bicgstab_lu_precond(h[0], h[1], h[2], h[3], maxit=2000, tol=1e-10)
                wall = time.time() - t0
                relerr = float(np.linalg.norm(xr - xt.cpu().numpy()) / np.linalg.norm(xt.cpu().numpy()))
                out["reference_gpu_ilu0"] = {"iterations": info["iterations"], "t_loop_s": dtr, "iters_per_s": info["iterations"] / max(dtr, 1e-9),
                                             "wall_s": wall, "rel_err_vs_xtrue": relerr,
                                             "what": "reference bicgstab_lu_precond (cuSPARSE SpMV/SpSV/csrilu02 + cuBLAS L1) on this GPU, tol 1e-10"}
                del xr, h
        except Exception as e:      # noqa: BLE001
            out["reference_gpu_ilu0"] = {"error": str(e)}

    # ---- same-box bar: modern cuSPARSE SpMV (cusparseSpMV through torch's CSR mat-vec) on the same matrix; library
    #      comparator only, never on the product path ----
    try:
        A = torch.sparse_csr_tensor(ia, ja, a, size=(n, n))
        y = A @ xt
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            y = A @ xt
        e1.record()
        torch.cuda.synchronize()
        cms = e0.elapsed_time(e1) / 20
        e0.record()
        for _ in range(20):
            s.spmv(xt.data_ptr(), y.data_ptr(), variant=1)
        e1.record()
        torch.cuda.synchronize()
        oms = e0.elapsed_time(e1) / 20
        e0.record()
        for _ in range(20):
            s.spmv(xt.data_ptr(), y.data_ptr())
        e1.record()
        torch.cuda.synchronize()
        ams = e0.elapsed_time(e1) / 20
        out["cusparse_spmv"] = {"cusparse_ms": cms, "cusparse_GBps": bytes_spmv(n, nnz) / cms / 1e6,
                                "ours_csr_ms": oms, "ours_csr_GBps": bytes_spmv(n, nnz) / oms / 1e6,
                                "ours_auto_ms": ams, "note": "plain y = A x, 20 back-to-back launches, CSR-algorithmic bytes; torch CSR mat-vec = cusparseSpMV"}
        del A, y
    except Exception as e:      # noqa: BLE001
        out["cusparse_spmv"] = {"error": str(e)}

    # ---- ILU0 mode on the same system (BASELINE config 3, second mode) ----
    if not args.no_ilu0:
        try:
            s2 = cm.Solver(n, stream=torch.cuda.current_stream().cuda_stream)
            s2.set_csr_device(nnz, a.data_ptr(), ia.data_ptr(), ja.data_ptr())
            sa = s2.analyze(cm.MODE_ILU0)
            st = s2.solve(cm.MODE_ILU0, b.data_ptr(), x.data_ptr(), maxit=5000, tol=1e-10)
            relerr = float(torch.linalg.norm(x - xt) / torch.linalg.norm(xt))
            out["ilu0"] = {"iterations": st["iterations"], "converged": bool(st["converged"]), "t_loop_s": st["t_loop"],
                           "iters_per_s": st["iterations"] / st["t_loop"], "t_analysis_s": sa["t_analysis"], "t_ilu0_s": sa["t_ilu0"],
                           "levels": [sa["levels_l"], sa["levels_u"]], "rel_err_vs_xtrue": relerr}
            s2.close()
            # the reference's live entry point end to end (bicgstab_lu_precond on malloc'ed host arrays, example.cpp:352): upload,
            # both analyses, ILU0, loop, download, release — next to reference_gpu_ilu0.wall_s of the reference's own GPU code
            hh = [t.cpu() for t in (a, ia, ja, b)]
            hx = torch.empty(n, dtype=torch.float64)
            walls = []
            for _ in range(2):
                wall, hst = host_solve(hh[0], hh[1], hh[2], hh[3], hx, mode=cm.MODE_ILU0)
                walls.append(wall)
            out["ilu0"]["e2e_host"] = {"wall_s": walls[-1], "wall_s_first_call": walls[0], "iterations": hst.iterations,
                                       "iters_per_s": hst.iterations / walls[-1], "t_h2d_s": hst.t_h2d, "t_analysis_s": hst.t_analysis,
                                       "t_ilu0_s": hst.t_ilu0, "t_loop_s": hst.t_loop, "t_d2h_s": hst.t_d2h,
                                       "call": "cudamat_bicgstab_host(MODE_ILU0, tol=1e-10), pageable host arrays, second call"}
            del hh, hx
            # opt-in multicolour ordering of the preconditioner: few levels, bandwidth-bound sweeps, weaker ILU(0)
            s3 = cm.Solver(n, stream=torch.cuda.current_stream().cuda_stream)
            s3.set_option("ilu0_reorder", 1)
            s3.set_csr_device(nnz, a.data_ptr(), ia.data_ptr(), ja.data_ptr())
            sa3 = s3.analyze(cm.MODE_ILU0)
            st3 = s3.solve(cm.MODE_ILU0, b.data_ptr(), x.data_ptr(), maxit=5000, tol=1e-10)
            out["ilu0_multicolor"] = {"iterations": st3["iterations"], "converged": bool(st3["converged"]), "t_loop_s": st3["t_loop"],
                                      "iters_per_s": st3["iterations"] / st3["t_loop"], "t_analysis_s": sa3["t_analysis"],
                                      "levels": [sa3["levels_l"], sa3["levels_u"]],
                                      "rel_err_vs_xtrue": float(torch.linalg.norm(x - xt) / torch.linalg.norm(xt))}
            s3.close()
        except Exception as e:      # noqa: BLE001
            out["ilu0"] = {"error": str(e)}

    # ---- BASELINE config 2: mat10000.mtx with ILU0 (example.cpp defaults: tol 1e-6, glibc-rand b) ----
    try:
        m, _, mia, mja, ma = cm.load_mm(os.path.join(ROOT, "tests", "golden", "mat10000.mtx"))
        rs = np.random.RandomState(0)
        bb = 1.0 + 4.0 * rs.rand(m)
        best = None
        wall_best = None
        for _ in range(3):
            t0 = time.perf_counter()
            xx, dt, st = cm.bicgstab_lu_precond(ma, mia, mja, bb, maxit=2000, tol=1e-6)
            w = time.perf_counter() - t0
            wall_best = w if wall_best is None else min(wall_best, w)
            if best is None or dt < best[0]:
                best = (dt, st)
        out["mat10000_ilu0"] = {"iterations": best[1]["iterations"], "t_loop_ms": best[0] * 1e3,
                                "wall_ms_host_call": wall_best * 1e3,     # upload, both analyses, ILU0, loop, download, release (pageable arrays)
                                "us_per_iteration": best[0] * 1e6 / max(best[1]["iterations"], 1),
                                "iters_per_s": best[1]["iterations"] / best[0], "roofline": "n/a (latency-bound)"}
    except Exception as e:      # noqa: BLE001
        out["mat10000_ilu0"] = {"error": str(e)}

    # ---- BASELINE config 4: generator.cpp-style random nonsymmetric diagonally dominant CSR, 50 M rows, irregular rows ----
    if not args.no_random:
        try:
            out["random_dd_50M"] = random_dd_run(cm, torch, args.random_rows)
        except Exception as e:      # noqa: BLE001
            out["random_dd_50M"] = {"error": str(e)}

    # ---- CPU baseline (reported, not the target): reference bicstab_omp on the host cores ----
    if not args.no_cpu:
        try:
            cb, _, _ = cpu_reference_run(N, budget_s=20.0)
            out["cpu_baseline"] = cb
        except Exception as e:      # noqa: BLE001
            out["cpu_baseline"] = {"value": None, "error": str(e)}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--grid", type=int, default=256, help="Poisson grid edge N (n = N^3 rows)")
    ap.add_argument("--chunk", type=int, default=250, help="iterations per restart inside the timed region")
    ap.add_argument("--variant", type=int, default=0, help="force an SpMV variant (1 rowlane, 2 staged)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--no-ilu0", action="store_true", help="skip the ILU0 extra")
    ap.add_argument("--no-refgpu", action="store_true", help="skip the run of the reference's own GPU code (oracle/_ref/libref_pbicgstab.so)")
    ap.add_argument("--no-random", action="store_true", help="skip the 50 M-row random matrix extra (BASELINE config 4)")
    ap.add_argument("--random-rows", type=int, default=50_000_000)
    ap.add_argument("--no-512", action="store_true", help="skip the 512^3 extra (BASELINE config 5)")
    ap.add_argument("--persist", type=int, default=-2, help="persistent cooperative iteration kernel: -1 auto, 0 off, 1 force (-2: library default)")
    ap.add_argument("--fuse", type=int, default=-1, help="MARCH loop: bit 0 folds the p update into SpMV 1, bit 1 the s update into SpMV 2 (-1: library default)")
    ap.add_argument("--no-csr", action="store_true", help="skip the second timed region with the plain CSR SpMV kernel")
    ap.add_argument("--no-converge", action="store_true", help="skip the full solve to 1e-10 (profiling runs)")
    ap.add_argument("--no-extras", action="store_true", help="skip e2e / ilu0 / mat10000 / cpu extras (profiling runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
