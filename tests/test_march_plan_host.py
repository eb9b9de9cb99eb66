"""CPU tests of the MARCH planner (pure host code exported as cudamat_march_plan_host): which superset patterns split into
planes -D / 0 / +D, and what the kernel is told about each entry.  No device needed."""


def test_poisson_7pt_planes(cm):
    for N, S, H in ((64, 2, 256), (128, 8, 256), (256, 32, 256), (512, 128, 512)):      # halo rounded up to 256 / 512
        off = [-N * N, -N, -1, 0, 1, N, N * N]
        p = cm.march_plan_host(off, [-1, -1, -1, 6, -1, -1, -1], N ** 3)
        assert p is not None, N
        assert p["D"] == N * N and p["S"] == S and p["P"] == N and p["H"] == H
        assert p["dz"] == [-1, 0, 0, 0, 0, 0, 1]
        assert p["loff"] == [0, -N, -1, 0, 1, N, 0]


def test_small_or_flat_grids_have_no_plan(cm):
    N = 32                                   # plane of 1024 rows: smaller than the 2048-row tile
    assert cm.march_plan_host([-N * N, -N, -1, 0, 1, N, N * N], None, N ** 3) is None
    N = 100                                  # 2-D 5-point stencil: a single plane, nothing to march along
    assert cm.march_plan_host([-N, -1, 0, 1, N], None, N * N) is None
    N = 96                                   # plane stride 9216 is not a multiple of the tile
    assert cm.march_plan_host([-N * N, -N, -1, 0, 1, N, N * N], None, N ** 3) is None


def test_rows_must_fill_whole_planes(cm):
    N = 128
    off = [-N * N, -N, -1, 0, 1, N, N * N]
    assert cm.march_plan_host(off, None, N ** 3 - 2048) is None
    assert cm.march_plan_host(off, None, N * N * 5) is not None          # N x N x 5 grid


def test_halo_limit_and_mixed_offsets(cm):
    D = 1 << 20
    # 8 entries, in-plane reach 300 -> halo rounded up to 512
    off = [-D - 300, -D, -7, 0, 7, 300, D, D + 2]
    p = cm.march_plan_host(off, None, 8 * D)
    assert p is not None and p["H"] == 512 and p["D"] == D
    assert p["dz"] == [-1, -1, 0, 0, 0, 0, 1, 1] and p["loff"] == [-300, 0, -7, 0, 7, 300, 0, 2]
    # an offset that is neither near 0 nor near +-D
    assert cm.march_plan_host([-D, 0, D // 2, D], None, 8 * D) is None
    # in-plane reach beyond the 512-element halo
    assert cm.march_plan_host([-D, -600, 0, 600, D], None, 8 * D) is None
