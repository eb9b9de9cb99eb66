"""Resource budget of the kernels of the default 256^3 loop, read from the built library with cuobjdump (no GPU needed).
A run-time branch added to k_spmv_march once cost the single-GPU MAKE_S kernel 156 bytes of spills and 30 % of its speed
(profiles/README.md, last session of round 2) without failing any test: the stack frames are pinned here."""
import os
import re
import shutil
import subprocess

import pytest

import __graft_entry__ as ge

CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"

# mangled-name fragment -> (max registers, max stack bytes); 2 CTAs x 512 threads per SM = 64 registers at most
BUDGET = {
    "k_spmv_marchILi0ELi1ELb0ELb0ELi1ELi1ELb0E": (64, 0),     # SpMV 1 + rhat.v           (LOAD_X, 1 dot, 7-point shape)
    "k_spmv_marchILi0ELi2ELb0ELb0ELi1ELi1ELb0E": (64, 0),     # SpMV 2 + t.s, t.t         (unfolded / ILU0 loop)
    "k_spmv_marchILi2ELi2ELb0ELb1ELi1ELi1ELb0E": (64, 24),    # MAKE_S: s update folded into SpMV 2 (single GPU)
    "k_update_xrILb0E": (64, 0),
    "k_update_xrILb1E": (64, 0),
    "k_update_pILb0E": (64, 0),
    "k_update_rx_ilu": (64, 0),
    "k_reduce_finish": (64, 16),              # the red[kMaxQ] array handed to apply_phase
}


@pytest.mark.skipif(not os.path.exists(CUOBJDUMP), reason="cuobjdump not installed")
def test_hot_loop_kernels_keep_their_register_and_stack_budget():
    lib = os.path.join(ge.PKG_DIR, "libcudamat_b200.so")
    if not os.path.exists(lib):
        ge.build()
    out = subprocess.run([CUOBJDUMP, "-res-usage", lib], capture_output=True, text=True, timeout=300).stdout
    usage = {}
    for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+)", out):
        usage[m.group(1)] = (int(m.group(2)), int(m.group(3)))
    assert usage, "cuobjdump printed no resource usage"
    for frag, (max_reg, max_stack) in BUDGET.items():
        hits = {k: v for k, v in usage.items() if frag in k}
        assert hits, "kernel %s not found in the library" % frag
        for name, (reg, stack) in hits.items():
            assert reg <= max_reg and stack <= max_stack, "%s: REG %d (<= %d) STACK %d (<= %d)" % (name, reg, max_reg, stack, max_stack)
