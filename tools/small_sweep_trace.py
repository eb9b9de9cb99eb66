#!/usr/bin/env python3
"""One L and one U sweep of the mat10000 ILU0 factor (single-CTA shared-memory sweep), timed with CUDA events; with
CUDAMAT_SWEEP_DEBUG=1 the library prints the SM cycles of every level."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
import numpy as np, torch
cm = ge.load_package()
for nm in (sys.argv[1:] or ["mat10000"]):
    m, _, ia, ja, a = cm.load_mm(os.path.join(ROOT, "tests", "golden", nm + ".mtx"))
    dev = torch.device("cuda:0")
    d_ia = torch.from_numpy(np.asarray(ia, dtype=np.int32) - int(ia[0])).to(dev); d_ja = torch.from_numpy(np.asarray(ja, dtype=np.int32) - int(ia[0])).to(dev)
    d_a = torch.from_numpy(np.asarray(a, dtype=np.float64)).to(dev)
    s = cm.Solver(m)
    s.set_csr_device(len(a), d_a.data_ptr(), d_ia.data_ptr(), d_ja.data_ptr())
    s.analyze(cm.MODE_ILU0)
    rhs = torch.rand(m, dtype=torch.float64, device=dev) + 1.0
    t = torch.empty_like(rhs); o = torch.empty_like(rhs)
    for _ in range(3):
        s.sptrsv(0, rhs.data_ptr(), t.data_ptr()); s.sptrsv(1, t.data_ptr(), o.data_ptr())
    torch.cuda.synchronize()
    if os.environ.get("CUDAMAT_SWEEP_DEBUG"):
        continue
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    best = [1e9, 1e9]
    for _ in range(20):
        e[0].record(); s.sptrsv(0, rhs.data_ptr(), t.data_ptr()); e[1].record(); s.sptrsv(1, t.data_ptr(), o.data_ptr()); e[2].record()
        torch.cuda.synchronize()
        best = [min(best[0], e[0].elapsed_time(e[1])), min(best[1], e[1].elapsed_time(e[2]))]
    print("%s: L %.1f us, U %.1f us" % (nm, best[0] * 1e3, best[1] * 1e3))
