#!/usr/bin/env python3
"""Static evidence from the built library (no GPU): per kernel of the loops, registers / stack (cuobjdump -res-usage) and counts
of the SASS mnemonics that show how it moves data — 128-bit global loads, LDGSTS (cp.async), UBLKCP (TMA bulk copy), SYNCS
(mbarrier), shared-memory traffic, fp64 arithmetic, shuffles, barriers — and that no tensor-core instruction exists.
usage: tools/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "cuda-mat_b200", "libcudamat_b200.so")
CUOBJDUMP = "/usr/local/cuda/bin/cuobjdump"
HOT = ["k_spmv_marchILi0ELi1ELb0ELb0ELi1ELi1ELb0E", "k_spmv_marchILi0ELi2ELb0ELb0ELi1ELi1ELb0E", "k_spmv_marchILi2ELi2ELb0ELb1ELi1ELi1ELb0E",
       "k_update_xrILb0E", "k_update_xrILb1E", "k_update_pILb0E", "k_update_sILb", "k_update_rx_ilu", "k_init_resid", "k_reduce_finish",
       "k_spmv_rowlaneILb0ELi1E", "k_spmv_stagedILb0ELi1E", "k_spmv_tiledILb0ELi1ELb0E", "k_spmv_streamILb0ELi1ELi0E", "k_spmv_stream_blkILb0ELi0ELi0E",
       "k_sptrsv_blockedILb0E", "k_sptrsv_blockedILb1E", "k_sptrsv_ringILb0E", "k_sptrsv_syncfreeILb0E", "k_ilu0_level", "k_bicgstab_persist"]
KEYS = [("LDG.E.128", r"\bLDG\.E\.128"), ("LDG (all)", r"\bLDG\."), ("STG.E.128", r"\bSTG\.E\.128"), ("STG (all)", r"\bSTG\."),
        ("LDGSTS", r"\bLDGSTS"), ("UBLKCP", r"\bUBLKCP"), ("SYNCS", r"\bSYNCS"), ("LDS", r"\bLDS"), ("STS", r"\bSTS"),
        ("DFMA", r"\bDFMA"), ("DADD", r"\bDADD"), ("DMUL", r"\bDMUL"), ("SHFL", r"\bSHFL"), ("BAR", r"\bBAR\."), ("ATOM/RED", r"\b(ATOM|RED)\b|\bATOMG|\bREDG"),
        ("LDL/STL (spill)", r"\b(LDL|STL)\b"), ("tensor (HMMA/IMMA/DMMA/UTCMMA/QGMMA)", r"\b(HMMA|IMMA|DMMA|BMMA|UTC\w*MMA|QGMMA|HGMMA)")]
res = subprocess.run([CUOBJDUMP, "-res-usage", LIB], capture_output=True, text=True).stdout
usage = {m.group(1): (m.group(2), m.group(3), m.group(4)) for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+)", res)}
sass = subprocess.run([CUOBJDUMP, "-sass", LIB], capture_output=True, text=True).stdout
archs = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
funcs = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1); funcs[cur] = []
    elif cur is not None and "/*" in line and ";" in line:
        funcs[cur].append(line)
print("# SASS / resource summary of cuda-mat_b200/libcudamat_b200.so (tools/sass_summary.py; cuobjdump of CUDA 12.9)")
print("# cubin architectures in the library:", ", ".join(archs), "| kernels:", len(funcs))
tot_tensor = 0
for name, body in funcs.items():
    tot_tensor += sum(1 for l in body if re.search(KEYS[-1][1], l))
print("# tensor-core instructions in the WHOLE library:", tot_tensor, "(the path is HBM-bound fp64 streaming work; deliberately none)")
print("# UBLKCP (TMA bulk copies) in the whole library:", sum(1 for b in funcs.values() for l in b if re.search(r"\bUBLKCP", l)),
      "| LDGSTS (cp.async):", sum(1 for b in funcs.values() for l in b if re.search(r"\bLDGSTS", l)),
      "| SYNCS (mbarrier):", sum(1 for b in funcs.values() for l in b if re.search(r"\bSYNCS", l)))
print()
hdr = ["kernel", "REG", "STACK", "SMEM(static)", "instr"] + [k for k, _ in KEYS]
print(" | ".join(hdr))
for frag in HOT:
    for name, body in funcs.items():
        if frag in name:
            u = usage.get(name, ("?", "?", "?"))
            short = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip().split("(")[0].replace("void cudamat::", "")
            row = [short, u[0], u[1], u[2], str(len(body))] + [str(sum(1 for l in body if re.search(p, l))) for _, p in KEYS]
            print(" | ".join(row))
