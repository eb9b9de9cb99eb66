"""CPU check of the reduction spec used by the fused dot epilogues: the packed butterfly that
csrc/internal.cuh `packed_pair` / `packed_butterfly` implement must give, for each of the M values, exactly the bits of one
plain butterfly (xor 16, 8, 4, 2, 1; every lane adds its partner's value to its own), and the lane that ends up holding
value k must be the one the kernels store from (k = (lane>>4 & 1) + 2 (lane>>3 & 1) + 4 (lane>>2 & 1))."""
import numpy as np
import pytest


def plain_butterfly(v):
    v = v.copy()
    lanes = np.arange(32)
    for s in (16, 8, 4, 2, 1):
        v = v + v[lanes ^ s]
    return v


def packed_pair(a, b, s):
    lanes = np.arange(32)
    up = (lanes & s) != 0
    send = np.where(up, a, b)
    keep = np.where(up, b, a)
    return keep + send[lanes ^ s]


def packed(vals):
    m = len(vals)
    v = [x.copy() for x in vals]
    lanes = np.arange(32)
    s = 16
    while m > 1:
        v = [packed_pair(v[2 * k], v[2 * k + 1], s) for k in range(m // 2)]
        m //= 2
        s //= 2
    z = v[0]
    while s >= 1:
        z = z + z[lanes ^ s]
        s //= 2
    return z


@pytest.mark.parametrize("m", [4, 8])
def test_packed_butterfly_is_bit_identical(m):
    rng = np.random.default_rng(7 + m)
    lanes = np.arange(32)
    for trial in range(200):
        scale = 10.0 ** rng.integers(-8, 9, size=(m, 32))
        vals = [rng.standard_normal(32) * scale[k] for k in range(m)]
        if trial == 0:
            vals[0][:] = 0.0
            vals[1][3] = -0.0
        z = packed(vals)
        idx = ((lanes >> 4) & 1) + 2 * ((lanes >> 3) & 1) + (4 * ((lanes >> 2) & 1) if m == 8 else 0)
        for k in range(m):
            want = plain_butterfly(vals[k])
            assert len(set(want.view(np.uint64).tolist())) == 1          # every lane of a plain butterfly agrees
            got = z[idx == k]
            assert len(got) == 32 // m
            assert np.array_equal(got.view(np.uint64), np.full(len(got), want.view(np.uint64)[0], dtype=np.uint64)), (m, k)
