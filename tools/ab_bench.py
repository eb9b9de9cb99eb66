#!/usr/bin/env python3
"""Same-box A/B of tuning builds (GPU only).  Build variants here with
    make -C cuda-mat_b200 alt ALT=<name> ALTFLAGS="-D..."        -> cuda-mat_b200/libcudamat_b200_<name>.so
(or from another commit: git stash / checkout, make alt ALT=prev, come back), then on the box
    python tools/ab_bench.py [--parity] [--repeat 2] [--args "--configs 5:0 --iters 40"] prev new ...
runs tools/spmv_bench.py once per library ('' or 'main' = the default build) in separate processes, optionally the GPU
parity tests against each library first, and prints one table."""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def lib_path(name):
    if name in ("", "main"):
        return os.path.join(ROOT, "cuda-mat_b200", "libcudamat_b200.so")
    return os.path.join(ROOT, "cuda-mat_b200", "libcudamat_b200_%s.so" % name)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("libs", nargs="+")
    ap.add_argument("--args", default="--configs 5:0 --iters 40")
    ap.add_argument("--repeat", type=int, default=1)
    ap.add_argument("--parity", action="store_true", help="run tests/test_gpu_parity.py against every library first")
    a = ap.parse_args()
    rows = []
    for name in a.libs:
        env = dict(os.environ, CUDAMAT_LIB=lib_path(name))
        if not os.path.exists(env["CUDAMAT_LIB"]):
            print("missing", env["CUDAMAT_LIB"]); continue
        if a.parity:
            r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-x", "-q"],
                               env=env, capture_output=True, text=True, timeout=600)
            print("%-10s parity: %s" % (name or "main", r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:]))
        for rep in range(a.repeat):
            r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "spmv_bench.py")] + a.args.split(), env=env,
                               capture_output=True, text=True, timeout=600)
            for line in r.stdout.splitlines():
                if line.startswith("{"):
                    d = json.loads(line); d["_lib"] = name or "main"; rows.append(d)
            if r.returncode:
                print("%-10s bench failed: %s" % (name or "main", r.stderr[-400:]))
    print("%-10s %7s %12s %12s %12s %10s" % ("lib", "variant", "plain_ms", "loop_spmv_ms", "ms_per_iter", "it/s"))
    for d in rows:
        print("%-10s %7d %12.4f %12.4f %12.4f %10.1f" % (d["_lib"], d["variant"], d["plain_spmv_ms"], d["loop_spmv_ms"], d["ms_per_iter"], d["iters_per_s"]))


if __name__ == "__main__":
    main()
