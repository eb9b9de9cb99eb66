"""CPU tests of the N > 1 path (world_size 2, gloo): the product's host-side planners
(cudamat_partition_rows / cudamat_halo_plan_host, the pure-host part of cudamat_comm_init) drive a
row-sharded BiCGSTAB whose per-shard arithmetic is played by the oracle primitives.  What must hold — and
what the GPU path relies on — is that sharding changes NOTHING: halo-remapped local SpMVs and the
zero-padded allreduce of tile partials reproduce the single-rank oracle bit for bit."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, N, port, out_dir):
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    cm, O = ge.load_package(), ge.load_oracle()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = N ** 3
    row0, row1 = cm.partition_rows(n, world, rank)
    nloc = row1 - row0
    ranges = [None] * world
    dist.all_gather_object(ranges, (row0, row1))
    row_starts = [r[0] for r in ranges] + [ranges[-1][1]]
    assert row_starts[0] == 0 and row_starts[-1] == n and all(r % 2048 == 0 for r in row_starts[:-1])
    ia, ja, a = O.poisson3d(N, row0, row1)                       # shard rows, GLOBAL columns
    halo, recv_cnt = cm.halo_plan_host(row0, row1, ja, row_starts)
    # what comm_init does with NCCL: tell every owner which of its rows this rank needs
    wanted = [None] * world
    dist.all_gather_object(wanted, halo.tolist())
    send_idx = []                                                # per peer: my local rows it wants
    for q in range(world):
        send_idx.append(np.array([c - row0 for c in wanted[q] if row0 <= c < row1], dtype=np.int64))
    recv_off = np.concatenate([[0], np.cumsum(recv_cnt)])
    assert recv_off[-1] == len(halo) and recv_cnt[rank] == 0
    # local / halo column numbering (k_remap_cols)
    jl = np.where((ja >= row0) & (ja < row1), ja - row0, nloc + np.searchsorted(halo, ja)).astype(np.int32)
    ntile_g = (n + 2047) // 2048
    tile0 = row0 // 2048

    def halo_exchange(vec):                                      # vec has nloc + nhalo entries
        bufs = [None] * world
        dist.all_gather_object(bufs, [vec[send_idx[q]].copy() for q in range(world)])
        for p in range(world):
            if recv_cnt[p]:
                vec[nloc + recv_off[p]: nloc + recv_off[p + 1]] = bufs[p][rank]

    def spmv(x_ext):
        halo_exchange(x_ext)
        return O.spmv(ia, jl, a, x_ext)

    def gdot(u, v):                                              # zero-padded allreduce of tile partials
        part = np.zeros(ntile_g)
        if nloc:
            t = O.dot_tiles(u[:nloc], v[:nloc])
            part[tile0: tile0 + len(t)] = t
        tt = torch.from_numpy(part)
        dist.all_reduce(tt)
        return O.combine_tiles(tt.numpy())

    ext = nloc + len(halo)
    xt = O.xtrue(1234, row0, nloc)
    xe = np.zeros(ext); xe[:nloc] = xt
    b = spmv(xe)
    # unpreconditioned loop, same operation order as oracle/oracle.c orc_bicgstab_unprec
    x0 = np.zeros(ext); x0[:nloc] = 1.0
    r = b - spmv(x0)
    r0 = r.copy()
    p = np.zeros(ext); v = np.zeros(nloc); s = np.zeros(ext)
    rho = alpha = omega = 1.0
    norm0 = np.sqrt(gdot(r, r))
    x = np.zeros(nloc)
    iters, tol = 0, 1e-10
    for it in range(5000):
        rho_ = gdot(r0, r)
        beta = (rho_ / rho) * (alpha / omega)
        q = (-omega) * v
        q = p[:nloc] + q
        q = beta * q
        p[:nloc] = r + q
        v = spmv(p)
        alpha = rho_ / gdot(r0, v)
        h = x0[:nloc] + alpha * p[:nloc]
        s[:nloc] = r + (-alpha) * v
        t = spmv(s)
        omega = gdot(t, s) / gdot(t, t)
        x = h + omega * s[:nloc]
        r = s[:nloc] + (-omega) * t
        iters = it + 1
        if np.sqrt(gdot(r, r)) < tol * norm0:
            break
        x0[:nloc] = x
        rho = rho_
    np.save(os.path.join(out_dir, "x_%d.npy" % rank), x)
    np.save(os.path.join(out_dir, "b_%d.npy" % rank), b)
    np.save(os.path.join(out_dir, "meta_%d.npy" % rank), np.array([row0, row1, iters, len(halo)]))
    dist.destroy_process_group()


@pytest.mark.parametrize("N", [16, 20])
def test_sharded_bicgstab_bit_identical(tmp_path, O, cm, N):
    world = 2
    port = 29500 + (os.getpid() % 2000) + N
    mp.spawn(_worker, args=(world, N, port, str(tmp_path)), nprocs=world, join=True)
    n = N ** 3
    ia, ja, a = O.poisson3d(N)
    xt = O.xtrue(1234, 0, n)
    b_ref = O.spmv(ia, ja, a, xt)
    x_ref, st = O.bicgstab_unprec(ia, ja, a, b_ref, maxit=5000, tol=1e-10)
    xs, bs, its = [], [], set()
    for rnk in range(world):
        meta = np.load(tmp_path / ("meta_%d.npy" % rnk))
        xs.append(np.load(tmp_path / ("x_%d.npy" % rnk)))
        bs.append(np.load(tmp_path / ("b_%d.npy" % rnk)))
        its.add(int(meta[2]))
        assert meta[3] == (N * N if meta[1] > meta[0] else 0)       # one halo plane per neighbour
    assert np.array_equal(np.concatenate(bs), b_ref)
    assert its == {st["iterations"]}
    assert np.array_equal(np.concatenate(xs), x_ref)


def test_partition_and_halo_planners(cm, O):
    # alignment rules: groups (2 Mi rows) when every rank can own one, else tiles (2048 rows)
    for n, world in ((256 ** 3, 8), (512 ** 3, 8), (256 ** 3, 2), (64 ** 3, 4), (5000, 3), (100, 2)):
        prev = 0
        for rank in range(world):
            r0, r1 = cm.partition_rows(n, world, rank)
            assert r0 == prev and r0 <= r1 <= n
            gran = 2048 * 1024 if n >= 2048 * 1024 * world else 2048
            assert r0 % gran == 0 or r0 == n
            prev = r1
        assert prev == n
    assert cm.partition_rows(256 ** 3, 8, 3) == (3 * 2 ** 21, 4 * 2 ** 21)
    # halo plan of the middle shard of a 3-way split: one plane from each neighbour, sorted, unique
    N = 16
    n = N ** 3
    starts = [0, 2048, 4096 - 2048 + 2048, n]
    starts = [0, 1024 * 2, 2048 + 1024, n]
    ia, ja, a = O.poisson3d(N, starts[1], starts[2])
    halo, cnt = cm.halo_plan_host(starts[1], starts[2], ja, starts)
    assert np.all(np.diff(halo) > 0) and cnt.tolist() == [N * N, 0, N * N]
    assert halo[0] == starts[1] - N * N and halo[-1] == starts[2] + N * N - 1
    with pytest.raises(cm.CudamatError):
        cm.halo_plan_host(0, 10, np.array([50], dtype=np.int32), [0, 10, 20])
