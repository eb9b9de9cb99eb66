"""CPU tests of the TILED SpMV host planner (csrc/rowclass.cu tiled_plan_host through the C ABI, no device needed):
x windows per 2048-row tile, shared-memory index of every entry, eligibility of the classes, superset pattern + masks.
The row classes are derived here from the oracle's matrices exactly as the device analysis defines them
(class = (length, column offsets ja - row [, values]), numbered by first row)."""
import numpy as np

KTILE = 2048


def classes_of(ia, ja, a, with_vals=True):
    n = len(ia) - 1
    seen, lens, offs, vals, hist = {}, [], [], [], []
    for i in range(n):
        lo, hi = ia[i], ia[i + 1]
        key = (tuple((ja[lo:hi] - i).tolist()), tuple(a[lo:hi].tolist()) if with_vals else ())
        c = seen.get(key)
        if c is None:
            c = seen[key] = len(lens)
            lens.append(hi - lo); offs.append(list(key[0])); vals.append(list(a[lo:hi])); hist.append(0)
        hist[c] += 1
    return lens, offs, vals, hist


def poisson_classes(N):
    """the 27 row classes of the 7-point Dirichlet stencil on an N^3 grid, without building the matrix"""
    lens, offs, vals, hist = [], [], [], []
    cnt = {"lo": 1, "mid": N - 2, "hi": 1}
    for bk in ("mid", "lo", "hi"):
        for bj in ("mid", "lo", "hi"):
            for bi in ("mid", "lo", "hi"):
                o = []
                if bk != "lo": o.append(-N * N)
                if bj != "lo": o.append(-N)
                if bi != "lo": o.append(-1)
                o.append(0)
                if bi != "hi": o.append(1)
                if bj != "hi": o.append(N)
                if bk != "hi": o.append(N * N)
                lens.append(len(o)); offs.append(o); vals.append([6.0 if x == 0 else -1.0 for x in o])
                hist.append(cnt[bk] * cnt[bj] * cnt[bi])
    return lens, offs, vals, hist


def test_poisson3d_plan(cm, O):
    N = 128                                                  # lines of 128, planes of 16384: {-N^2}, {-N..N}, {+N^2}
    lens, offs, vals, hist = poisson_classes(N)
    assert sum(hist) == N ** 3
    p = cm.tiled_plan_host(lens, offs, vals, hist, N ** 3)
    assert p is not None
    assert [w[0] for w in p["windows"]] == [-N * N, -N, N * N]
    assert [w[1] for w in p["windows"]] == [KTILE + 2, KTILE + 2 * N + 2, KTILE + 2]          # even lengths
    bases = [w[2] for w in p["windows"]]
    assert bases == [0, KTILE + 2, 2 * KTILE + 2 * N + 4]
    assert p["ok_mask"] == (1 << 27) - 1                                                    # every class fits
    assert p["smem_bytes"] == 8 * (3 * KTILE + 2 * N + 6) + 208 * 27
    # shared-memory index of an entry = window base + (offset - first offset of the window)
    for c in range(27):
        for q, o in enumerate(offs[c]):
            g = 0 if o == -N * N else 2 if o == N * N else 1
            assert p["disp"][c, q] == bases[g] + (o - p["windows"][g][0]), (c, q)
    # superset pattern: the 7-point stencil, values 6 / -1, byte offsets of its entries
    assert p["sup_len"] == 7
    want_off = [-N * N, -N, -1, 0, 1, N, N * N]
    assert p["sup_val"] == [-1.0, -1.0, -1.0, 6.0, -1.0, -1.0, -1.0]
    for k, o in enumerate(want_off):
        g = 0 if o == -N * N else 2 if o == N * N else 1
        assert p["sup_boff"][k] == 8 * (bases[g] + o - p["windows"][g][0])
    for c in range(27):
        assert p["class_mask"][c] == sum(1 << want_off.index(o) for o in offs[c]), c
    assert p["class_mask"][0] == 0x7f                       # the interior class holds the whole pattern


def test_small_grid_merges_the_windows(cm, O):
    """N = 24: the plane distance (576) is below the clustering gap, one window covers all seven offsets; the classes
    come from the oracle's matrix, extracted the way the device analysis defines them"""
    N = 24
    ia, ja, a = O.poisson3d(N)
    lens, offs, vals, hist = classes_of(ia, ja, a)
    assert len(lens) == 27 and sorted(hist) == sorted(poisson_classes(N)[3])
    p = cm.tiled_plan_host(lens, offs, vals, hist, N ** 3)
    assert p is not None and p["windows"] == [(-N * N, KTILE + 2 * N * N + 2, 0)]
    assert p["sup_len"] == 7 and p["ok_mask"] == (1 << 27) - 1
    assert p["sup_boff"] == [8 * (o + N * N) for o in (-N * N, -N, -1, 0, 1, N, N * N)]


def test_value_conflict_disables_pattern_but_not_the_plan(cm, O):
    ia, ja, a = O.poisson3d(16)
    a = a.copy()
    for i in range(len(ia) - 1):
        lo, hi = ia[i], ia[i + 1]
        a[lo:hi][ja[lo:hi] == i] = hi - lo - 0.5            # diagonal depends on the row's degree
    lens, offs, vals, hist = classes_of(ia, ja, a)
    p = cm.tiled_plan_host(lens, offs, vals, hist, 16 ** 3)
    assert p is not None and p["sup_len"] == 0 and p["ok_mask"] == (1 << len(lens)) - 1
    # offsets-only dictionary of the same matrix: the pattern exists again (values come from CSR)
    lens0, offs0, _, hist0 = classes_of(ia, ja, a, with_vals=False)
    p0 = cm.tiled_plan_host(lens0, offs0, None, hist0, 16 ** 3, with_vals=False)
    assert p0 is not None and p0["sup_len"] == 7


def test_more_than_eight_offsets_has_no_pattern(cm):
    N = 64                                                   # 2-D 9-point stencil, interior class only + one edge class
    inter = [-N - 1, -N, -N + 1, -1, 0, 1, N - 1, N, N + 1]
    edge = [-N, -N + 1, 0, 1, N, N + 1]
    p = cm.tiled_plan_host([9, 6], [inter, edge], [[1.0] * 9, [1.0] * 6], [3844, 124], N * N)
    assert p is not None and len(p["windows"]) == 1 and p["sup_len"] == 0
    assert p["windows"][0][0] == -N - 1 - 1                 # first offset rounded down to even
    assert p["ok_mask"] == 3


def test_rare_far_class_is_left_to_the_gather_path(cm):
    n = 1 << 20
    inter = [-1, 0, 1]
    wrap = [-1, 0, 1, n - 4096]                              # rare class with a far column: not in the windows
    p = cm.tiled_plan_host([3, 4], [inter, wrap], [[-1.0, 2.0, -1.0], [-1.0, 2.0, -1.0, 0.5]], [n - 8, 8], n)
    assert p is not None and len(p["windows"]) == 1
    assert p["ok_mask"] == 1                                 # tiles holding a `wrap` row take the gather path
    assert p["sup_len"] == 3 and p["class_mask"][0] == 7 and p["class_mask"][1] == 0


def test_no_plan_when_windows_do_not_fit(cm):
    n = 1 << 22
    # five far-apart offsets in the frequent class: more than 4 windows
    offs = [-3 * 65536, -65536, 0, 65536, 3 * 65536]
    assert cm.tiled_plan_host([5], [offs], [[1.0] * 5], [n], n) is None
    # one window wider than 100 KB of shared memory (offset gaps of <= 4096 chain into a single cluster)
    chain = list(range(0, 16 * 4000, 4000))
    assert cm.tiled_plan_host([16], [chain], [[1.0] * 16], [n], n) is None


def test_unsorted_offsets_disable_pattern(cm):
    p = cm.tiled_plan_host([3], [[1, 0, -1]], [[1.0, 2.0, 3.0]], [4096], 4096)
    assert p is not None and p["sup_len"] == 0 and p["ok_mask"] == 1
