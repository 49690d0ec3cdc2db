"""world_size-2 gloo test (CPU) of the gene-sharded protocol: shard bounds, the cross-shard sums the
library all-reduces (per-cell drop-out partial log-likelihoods, EM log-likelihood), the global
rowMeans minimum and the final gather of per-gene results.  The per-shard arithmetic is done by the
CPU oracle here (test infrastructure); on GPUs the same exchanges are made by liblpb200.so through
Engine.set_collective (lineagepulse_b200/distributed.py)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    from lineagepulse_b200 import _cabi as A
    from lineagepulse_b200.distributed import allreduce_host, gather_genes, global_min, shard_bounds
    from tests.helpers import design_for, make_dataset, model
    from tests.oracle_binding import Oracle
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        oracle = Oracle(nthreads=2)
        ds = make_dataset(60, 14, 8, 8, seed=3)
        X, G, N = ds["X"], ds["G"], ds["N"]
        b = shard_bounds(G, world)
        lo, hi = b[rank], b[rank + 1]
        shard = dict(ds, X=np.ascontiguousarray(X[lo:hi]), G=hi - lo)
        d_full, d_loc = design_for(ds), design_for(shard)
        m = model("constant", drop="logistic")
        # global quantities from per-shard pieces
        gmin = global_min(float(oracle.row_means(shard["X"], d_loc).min()))
        p_loc = oracle.init_params(shard["X"], d_loc, m)
        p_loc.mat_drop[:, 0] = -1.3
        ll_loc = oracle.eval_loglik(shard["X"], d_loc, m, p_loc)
        ll_sum = allreduce_host(np.array([ll_loc.sum()]))[0]            # EM log-likelihood (fitModel.R:328)
        # per-cell partial sums over the shard's genes, summed across shards (the fitPi exchange)
        part = np.zeros(N)
        for j in range(N):
            pi = oracle.dropout_rate([p_loc.mat_drop[j, 0]], [1.0])
            for i in range(hi - lo):
                part[j] += oracle.ll_obs(shard["X"][i, j], p_loc.mat_mu[i, 0], p_loc.mat_disp[i, 0], pi)
        cell_ll = allreduce_host(part)
        ll_all = gather_genes(ll_loc)
        if rank == 0:
            p_full = oracle.init_params(X, d_full, m)
            p_full.mat_drop[:, 0] = -1.3
            ll_full = oracle.eval_loglik(X, d_full, m, p_full)
            ref_cell = np.zeros(N)
            for j in range(N):
                pi = oracle.dropout_rate([-1.3], [1.0])
                for i in range(G):
                    ref_cell[j] += oracle.ll_obs(X[i, j], p_full.mat_mu[i, 0], p_full.mat_disp[i, 0], pi)
            out.put({"bounds": b, "gmin": gmin, "gmin_ref": float(oracle.row_means(X, d_full).min()),
                     "ll_sum": ll_sum, "ll_sum_ref": float(ll_full.sum()), "ll_all": ll_all, "ll_full": ll_full,
                     "cell": cell_ll, "cell_ref": ref_cell})
    finally:
        dist.destroy_process_group()


def test_gene_sharded_exchanges_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = out.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res["bounds"][0] == 0 and res["bounds"][-1] == len(res["ll_full"])
    assert res["gmin"] == res["gmin_ref"]
    assert abs(res["ll_sum"] - res["ll_sum_ref"]) <= 1e-12 * abs(res["ll_sum_ref"])
    assert np.array_equal(res["ll_all"], res["ll_full"])                 # gather restores gene order
    assert np.allclose(res["cell"], res["cell_ref"], rtol=1e-13, atol=0)  # sharded sums == unsharded


def test_shard_bounds():
    from lineagepulse_b200.distributed import shard_bounds
    assert shard_bounds(10, 4, align=1) == [0, 3, 6, 8, 10]
    b = shard_bounds(20000, 8)
    assert b[0] == 0 and b[-1] == 20000 and len(b) == 9
    assert all(x % 16 == 0 for x in b[:-1])          # shards start on summation-block boundaries
    assert max(np.diff(b)) - min(np.diff(b)) <= 16
    assert shard_bounds(10, 4) == [0, 10, 10, 10, 10]  # fewer blocks than shards: trailing shards are empty


def _comm_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import ctypes as C
    import torch.distributed as dist
    from lineagepulse_b200 import _cabi as A
    from lineagepulse_b200.distributed import init_comm
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lib = A.load_library()

        class FakeEngine:   # the two calls init_comm makes; the real ones need a GPU (lpb200_comm_init = ncclCommInitRank)
            def comm_unique_id(self):
                buf = C.create_string_buffer(A.UNIQUE_ID_BYTES)
                assert lib.lpb200_comm_unique_id(buf) == 0      # the library's own NCCL binding (dlopen), no GPU needed
                return buf.raw

            def init_comm(self, r, w, uid):
                self.got = (r, w, uid)

        e = FakeEngine()
        r, w = init_comm(e)
        out.put((rank, r, w, e.got[0], e.got[1], e.got[2]))
    finally:
        dist.destroy_process_group()


def test_library_communicator_bootstrap_world2():
    """One process per GPU: rank 0 creates the NCCL unique id inside the library, torch.distributed only carries its
    128 bytes; every rank hands the same id with its own rank to lpb200_comm_init (distributed.init_comm)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_comm_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [(r[0], r[1], r[2], r[3], r[4]) for r in res] == [(0, 0, 2, 0, 2), (1, 1, 2, 1, 2)]
    assert res[0][5] == res[1][5] and len(res[0][5]) == 128 and any(res[0][5])
