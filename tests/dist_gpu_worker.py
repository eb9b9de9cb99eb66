"""torchrun worker for tests/test_gpu_multi.py: row-sharded BiCGSTAB over NCCL, one rank per GPU, checked
bit for bit against the single-GPU solve of the same system (which the other tests pin to the oracle)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402


def main():
    cm = ge.load_package()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    f64 = dict(dtype=torch.float64, device="cuda")
    grids = [int(v) for v in sys.argv[1].split(",")]
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    for N in grids:
        n = N ** 3
        row0, row1 = cm.partition_rows(n, world, rank)
        nloc = row1 - row0
        nnz = cm.poisson3d_nnz(N, row0, row1)
        ia = torch.empty(nloc + 1, dtype=torch.int32, device="cuda")
        ja = torch.empty(max(nnz, 1), dtype=torch.int32, device="cuda")
        a = torch.empty(max(nnz, 1), **f64)
        cm.gen_poisson3d_device(N, row0, row1, ia.data_ptr(), ja.data_ptr(), a.data_ptr())
        s = cm.Solver(n, row0, row1)
        s.set_csr_device(nnz, a.data_ptr(), ia.data_ptr(), ja.data_ptr())
        if rank == 0:
            uid = torch.tensor(list(cm.Comm.unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(uid, 0)
        cm.Comm.init(s, bytes(uid.cpu().tolist()), rank, world)
        p2p = cm.Comm.p2p_enabled(s)
        s.analyze(cm.MODE_PLAIN)
        xt = torch.empty(nloc, **f64)
        cm.gen_xtrue_device(1234, row0, nloc, xt.data_ptr())
        b = torch.empty(nloc, **f64)
        s.spmv(xt.data_ptr(), b.data_ptr())
        x = torch.zeros(nloc, **f64)
        st = s.solve(cm.MODE_PLAIN, b.data_ptr(), x.data_ptr(), maxit=5000, tol=1e-10)
        dt = s.dot(x.data_ptr(), b.data_ptr())
        torch.cuda.synchronize()
        # the same solve with the MARCH SpMV forced on the shard (neighbour planes read from the halo region) and the s update folded
        # into SpMV 2 (boundary planes pushed by k_update_s_boundary): must reproduce the bits of the default path
        march_ok = True
        if N >= 64 and p2p:
            s2 = cm.Solver(n, row0, row1)
            s2.set_option("march_shards", 2); s2.set_option("shard_fuse", 1); s2.set_option("persist", 0)
            s2.set_csr_device(nnz, a.data_ptr(), ia.data_ptr(), ja.data_ptr())
            cm.Comm.init(s2, bytes(uid.cpu().tolist()), rank, world)
            s2.analyze(cm.MODE_PLAIN)
            x2 = torch.zeros(nloc, **f64)
            st2 = s2.solve(cm.MODE_PLAIN, b.data_ptr(), x2.data_ptr(), maxit=5000, tol=1e-10)
            torch.cuda.synchronize()
            # the fold needs the peer-memory path on THIS handle (a handle whose IPC set-up failed solves over NCCL, unfolded)
            p2p2 = cm.Comm.p2p_enabled(s2)
            parts = (torch.equal(x2, x), st2["iterations"] == st["iterations"], st2["spmv_variant"] == 6, st2["fused"] == 2, p2p2)
            if not all(parts):
                print("DIST-MARCH-SHARD rank %d N=%d: x_equal=%s iterations %d/%d variant=%d fused=%d p2p=%d max|dx|=%.3e"
                      % (rank, N, parts[0], st2["iterations"], st["iterations"], st2["spmv_variant"], st2["fused"], int(p2p2),
                         float((x2 - x).abs().max())), flush=True)
            flag = torch.tensor([int(all(parts))], device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            march_ok = bool(flag.item())
            s2.close()
        # block-Jacobi ILU(0) on the same sharded handle (each rank factors its diagonal block): must converge to the
        # same solution; the iteration count differs from the global ILU(0) of a single GPU by construction
        s.analyze(cm.MODE_ILU0)
        xi = torch.zeros(nloc, **f64)
        sti = s.solve(cm.MODE_ILU0, b.data_ptr(), xi.data_ptr(), maxit=5000, tol=1e-10)
        e2 = torch.stack([torch.sum((xi - xt) ** 2), torch.sum(xt ** 2)])
        dist.all_reduce(e2)
        ilu_err = float(torch.sqrt(e2[0] / e2[1]))
        ilu_ok = bool(sti["converged"]) and ilu_err <= 1e-5 and sti["iterations"] < st["iterations"]
        # gather shards on every rank (equal-size padding)
        per = max(cm.partition_rows(n, world, r)[1] - cm.partition_rows(n, world, r)[0] for r in range(world))
        pad = torch.zeros(per, **f64); pad[:nloc] = x
        padb = torch.zeros(per, **f64); padb[:nloc] = b
        padi = torch.zeros(per, **f64); padi[:nloc] = xi
        xs = [torch.zeros(per, **f64) for _ in range(world)]
        bs = [torch.zeros(per, **f64) for _ in range(world)]
        xis = [torch.zeros(per, **f64) for _ in range(world)]
        dist.all_gather(xs, pad); dist.all_gather(bs, padb); dist.all_gather(xis, padi)
        s.close()
        if rank == 0:
            sizes = [cm.partition_rows(n, world, r)[1] - cm.partition_rows(n, world, r)[0] for r in range(world)]
            xg = torch.cat([xs[r][:sizes[r]] for r in range(world)])
            bg = torch.cat([bs[r][:sizes[r]] for r in range(world)])
            # single-GPU reference run of the same global system
            nz = cm.poisson3d_nnz(N)
            ia1 = torch.empty(n + 1, dtype=torch.int32, device="cuda")
            ja1 = torch.empty(nz, dtype=torch.int32, device="cuda")
            a1 = torch.empty(nz, **f64)
            cm.gen_poisson3d_device(N, 0, n, ia1.data_ptr(), ja1.data_ptr(), a1.data_ptr())
            s1 = cm.Solver(n)
            s1.set_csr_device(nz, a1.data_ptr(), ia1.data_ptr(), ja1.data_ptr())
            s1.analyze(cm.MODE_PLAIN)
            xt1 = torch.empty(n, **f64); cm.gen_xtrue_device(1234, 0, n, xt1.data_ptr())
            b1 = torch.empty(n, **f64); s1.spmv(xt1.data_ptr(), b1.data_ptr())
            x1 = torch.zeros(n, **f64)
            st1 = s1.solve(cm.MODE_PLAIN, b1.data_ptr(), x1.data_ptr(), maxit=5000, tol=1e-10)
            d1 = s1.dot(x1.data_ptr(), b1.data_ptr())
            torch.cuda.synchronize()
            # block-Jacobi ILU(0) against the CPU oracle's restatement of it (oracle.c orc_bicgstab_ilu0_blocks: ILU(0) of the
            # diagonal blocks of the same row partition): same iteration count, same bits (sizes the oracle walks in seconds)
            blk_ok = True
            if n <= 64 ** 3:
                O = ge.load_oracle()
                oia, oja, oa = O.poisson3d(N)
                rs = [0] + [cm.partition_rows(n, world, r)[1] for r in range(world)]
                xo, so = O.bicgstab_ilu0_blocks(oia, oja, oa, bg.cpu().numpy(), rs, maxit=5000, tol=1e-10)
                xig = torch.cat([xis[r][:sizes[r]] for r in range(world)]).cpu().numpy()
                blk_ok = so["iterations"] == sti["iterations"] and bool(so["converged"]) and bool((xig == xo).all())
                print("DIST-BLOCK-ILU0 N=%d world=%d: iterations %d (oracle %d) x_equal=%s" % (N, world, sti["iterations"], so["iterations"], bool((xig == xo).all())), flush=True)
            # 256^3: the oracle's run takes minutes; its committed result (tests/golden/poisson256_blockilu0_oracle.json) is compared:
            # the iteration count must agree (it did on 2 and 8 GPUs in round 2), the digest of x is reported
            if N == 256:
                import hashlib, json
                gp = os.path.join(ROOT, "tests", "golden", "poisson256_blockilu0_oracle.json")
                ent = json.load(open(gp))["partitions"].get(str(world)) if os.path.exists(gp) else None
                if ent:
                    xig = torch.cat([xis[r][:sizes[r]] for r in range(world)]).cpu().numpy()
                    sha = hashlib.sha256(xig.tobytes()).hexdigest()
                    blk_ok = ent["iterations"] == sti["iterations"]
                    print("DIST-BLOCK-ILU0 N=256 world=%d: iterations %d (oracle %d) x_sha256_equal=%s"
                          % (world, sti["iterations"], ent["iterations"], sha == ent["x_sha256"]), flush=True)
            ok = (torch.equal(bg, b1) and torch.equal(xg, x1) and st["iterations"] == st1["iterations"]
                  and bool(st["converged"]) and dt == d1 and ilu_ok and march_ok and blk_ok)
            print("DIST N=%d world=%d p2p=%d variant=%d iters=%d/%d b_equal=%s x_equal=%s dot_equal=%s loop_ms=%.2f/%.2f block_ilu0_iters=%d err=%.1e march_fold_shard=%s %s"
                  % (N, world, int(p2p), st["spmv_variant"], st["iterations"], st1["iterations"], torch.equal(bg, b1), torch.equal(xg, x1), dt == d1,
                     st["t_loop"] * 1e3, st1["t_loop"] * 1e3, sti["iterations"], ilu_err, march_ok, "OK" if ok else "MISMATCH"), flush=True)
            s1.close()
        dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
