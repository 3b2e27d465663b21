"""Multi-GPU tests at world = 2, 4 and 8 (each runs when the box has that many GPUs: `gpurun --gpus N`; skipped on a
single-GPU box): NCCL item-sharded / grid-sharded scoring and the fused SpMM + peer-store all-gather (TMA bulk stores and
the SM-issued variant, sparse need masks, the two-range last-layer scatter, users-only replication, NVLS multicast) must
reproduce the single-GPU results — ids bit for bit, propagated rows within 1e-5 norm-wise."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import coldrec_oracle as O

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _case():
    rng = np.random.default_rng(77)
    n_users, n_items, n_q = 3000, 40000, 1500
    U = (rng.standard_normal((n_users, 64)) * 0.125).astype(np.float32)
    I = (rng.standard_normal((n_items, 64)) * 0.125).astype(np.float32)
    uids = rng.choice(n_users, n_q, replace=False).astype(np.int32)
    rows = [np.sort(rng.choice(n_items, int(rng.integers(0, 80)), replace=False)) for _ in range(n_q)]
    rowptr = np.zeros(n_q + 1, dtype=np.int64); np.cumsum([len(r) for r in rows], out=rowptr[1:])
    col = np.concatenate(rows).astype(np.int32)
    gt = [np.sort(rng.choice(n_items, 6, replace=False)) for _ in range(n_q)]
    gt_rowptr = np.arange(0, 6 * n_q + 1, 6, dtype=np.int64)
    gt_col = np.concatenate(gt).astype(np.int32)
    eu, ei = rng.integers(0, n_users, 200000), rng.integers(0, n_items, 200000)
    ei[:3000] = 7                       # one long row (> 512 nonzeros) so the split path crosses the partition too
    content = rng.standard_normal((n_items, 40)).astype(np.float32)
    gar = {"0.weight": (rng.standard_normal((128, 40)) * 0.2).astype(np.float32), "0.bias": np.zeros(128, np.float32),
           "2.weight": (rng.standard_normal((64, 128)) * 0.2).astype(np.float32), "2.bias": np.zeros(64, np.float32)}
    cold = np.sort(rng.choice(n_items, n_items // 5, replace=False)).astype(np.int64)
    return dict(U=U, I=I, uids=uids, rowptr=rowptr, col=col, gt_rowptr=gt_rowptr, gt_col=gt_col, eu=eu, ei=ei,
                n_users=n_users, n_items=n_items, content=content, gar=gar, cold=cold)


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from coldrec_b200 import ops
        from coldrec_b200.dist import RowPartitionedGraph, ShardedFullRankScorer, shard_range
        from coldrec_b200.scoring import EvalPlan
        c = _case()
        t = lambda a: torch.from_numpy(a).to(dev)
        plan = EvalPlan.from_arrays(t(c["uids"]), t(c["rowptr"]), t(c["col"]), t(c["gt_rowptr"]), t(c["gt_col"]))
        sc = ShardedFullRankScorer(20, ops.SCORE_TF32_CHECKED)
        b, e = shard_range(c["n_items"], rank, world)
        s, i = sc.topk(t(c["U"]), t(c["I"][b:e].copy()), b, plan)
        perf = sc.metrics(i, plan, [10, 20], rounded=False)
        lo, hi = sc.user_slice(plan.n_q)
        from coldrec_b200.dist import GridShardedFullRankScorer
        grid = {}
        for S in [d_ for d_ in (1, 2, 4, 8) if world % d_ == 0 and d_ <= world]:   # user-sharded ... grid (W=8: 4 x 2, 2 x 4) ... item-sharded
            gsc = GridShardedFullRankScorer(20, S, ops.SCORE_TF32_CHECKED)
            gb, ge = gsc.item_range(c["n_items"])
            g_s, g_i = gsc.topk(t(c["U"]), t(c["I"][gb:ge].copy()), gb, plan)
            grid[S] = (g_i.cpu().numpy(), gsc.user_slice(plan.n_q), gsc.metrics(g_i, plan, [10, 20], rounded=False))
        # generators over the item-sharded catalogue: each rank generates its rows, feeds its local sweep
        from coldrec_b200 import towers
        from coldrec_b200.dist import ShardedItemGenerator
        gst = {k: t(v) for k, v in c["gar"].items()}
        gen = ShardedItemGenerator(GridShardedFullRankScorer(20, world, ops.SCORE_TF32_CHECKED), c["n_items"])
        shard = gen.generate(lambda C: towers.gar_generate(gst, C), t(c["content"]))
        _, gen_i = gen.topk(t(c["U"]), shard, plan)
        base = gen.rows(t(c["I"])).clone()
        gen.overwrite_cold(lambda C, rows, out: towers.gar_generate(gst, C, rows=rows, out=out), base, t(c["content"]), t(c["cold"]))
        _, gen_i2 = gen.topk(t(c["U"]), base, plan)
        gen_out = dict(i=gen_i.cpu().numpy(), i2=gen_i2.cpu().numpy(), slice=gen.scorer.user_slice(plan.n_q))
        adj = O.normalize_graph_mat(O.bipartite_adjacency(c["eu"], c["ei"], c["n_users"], c["n_items"])).tocsr()
        adj.sort_indices()
        G = RowPartitionedGraph(adj.indptr.astype(np.int64), adj.indices.astype(np.int64), adj.data.astype(np.float32), dev,
                                segments=(c["n_users"], c["n_items"]))
        E0 = torch.cat([t(c["U"]), t(c["I"])])
        out_nccl = G.propagate(E0, 3).cpu().numpy()
        out_p2p = G.propagate_p2p(E0, 3).cpu().numpy()
        out_p2p_b = G.propagate_p2p(E0, 2, include_ego=False).cpu().numpy()     # buffers reused across calls
        out_p2p_c = G.from_padded(G.propagate_p2p(G.to_padded(E0), 3, padded_io=True)).cpu().numpy()   # padded numbering in and out
        out_p2p_d = G.propagate_p2p(E0, 1, copy=False).cpu().numpy()            # single layer: first == last; view of the result table
        out_p2p_e = G.propagate_p2p(E0, 3, sparse=False).cpu().numpy()          # dense all-gather (every row to every GPU)
        out_p2p_g = G.propagate_p2p(E0, 3, multicast=True).cpu().numpy()        # NVLS multimem.st where the node has it
        os.environ["CR_SPMM_PEER_ST"] = "1"                                     # round-1 path: SM-issued 16-byte peer stores
        out_p2p_h = G.propagate_p2p(E0, 3).cpu().numpy()
        del os.environ["CR_SPMM_PEER_ST"]
        os.environ["CR_SPMM_FORCE_BIG"] = "1"                                   # the bandwidth-bound geometry (16 lanes x float4, 128 rows per warp)
        out_p2p_i = G.propagate_p2p(E0, 3).cpu().numpy()
        del os.environ["CR_SPMM_FORCE_BIG"]
        # item-sharded consumers: user rows replicated, item rows stay with their owner
        out_p2p_f = G.propagate_p2p(E0, 3, replicate_result=(0,)).cpu().numpy()
        (ub, ue), (ib, ie) = G.parts[rank]
        ret[rank] = dict(s=s.cpu().numpy(), i=i.cpu().numpy(), lo=lo, hi=hi, perf=perf, nccl=out_nccl, p2p=out_p2p, p2p_b=out_p2p_b,
                         p2p_c=out_p2p_c, p2p_d=out_p2p_d, p2p_e=out_p2p_e, p2p_f=out_p2p_f, p2p_g=out_p2p_g, p2p_h=out_p2p_h, p2p_i=out_p2p_i, own_items=(ib, ie), has_multicast=G.has_multicast,
                         need_copies=G.need_copies, grid=grid, gen=gen_out)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(900)
@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_gpu_sharded_scoring_and_fused_allgather_propagation(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs (gpurun --gpus {world})")
    from coldrec_b200 import ops
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    c = _case()
    dev = torch.device("cuda", 0)
    t = lambda a: torch.from_numpy(a).to(dev)
    s1, i1, _ = ops.score_topk(t(c["U"]), t(c["I"]), 20, user_ids=t(c["uids"]), mask_rowptr=t(c["rowptr"]), mask_col=t(c["col"]),
                               precision=ops.SCORE_EXACT_F32)
    s1, i1 = s1.cpu().numpy(), i1.cpu().numpy()
    for r in range(world):
        lo, hi = ret[r]["lo"], ret[r]["hi"]
        assert np.array_equal(ret[r]["i"], i1[lo:hi]), "sharded ids must equal the single-GPU sweep bit for bit"
        assert np.allclose(ret[r]["s"], s1[lo:hi], atol=1e-6)
    want = O.metrics_from_topk(i1.astype(np.int64), c["gt_rowptr"], c["gt_col"].astype(np.int64), [10, 20])
    for r in range(world):
        assert sorted(ret[r]["grid"]) == [d_ for d_ in (1, 2, 4, 8) if world % d_ == 0]
        for S, (g_i, (glo, ghi), g_perf) in ret[r]["grid"].items():
            assert np.array_equal(g_i, i1[glo:ghi]) and np.allclose(g_perf, want, atol=1e-9), f"grid S={S} rank {r}"
        assert np.allclose(ret[r]["perf"], want, atol=1e-9)
    # sharded generators == generate the whole catalogue on one GPU, sweep once
    from coldrec_b200 import towers
    gst = {k: t(v) for k, v in c["gar"].items()}
    full_a = towers.gar_generate(gst, t(c["content"]))
    full_b = t(c["I"]).clone()
    towers.gar_generate(gst, t(c["content"]), rows=t(c["cold"]).to(torch.int32), out=full_b)
    kw = dict(user_ids=t(c["uids"]), mask_rowptr=t(c["rowptr"]), mask_col=t(c["col"]), precision=ops.SCORE_EXACT_F32)
    want_a = ops.score_topk(t(c["U"]), full_a, 20, **kw)[1].cpu().numpy()
    want_b = ops.score_topk(t(c["U"]), full_b, 20, **kw)[1].cpu().numpy()
    for r in range(world):
        lo, hi = ret[r]["gen"]["slice"]
        assert np.array_equal(ret[r]["gen"]["i"], want_a[lo:hi]), f"rank {r}: sharded whole-table generator + sweep"
        assert np.array_equal(ret[r]["gen"]["i2"], want_b[lo:hi]), f"rank {r}: sharded cold-row overwrite + sweep"
    adj = O.normalize_graph_mat(O.bipartite_adjacency(c["eu"], c["ei"], c["n_users"], c["n_items"]))
    Ut, It = torch.from_numpy(c["U"]), torch.from_numpy(c["I"])
    ref = torch.cat(O.propagate(adj, Ut, It, 3)).numpy()
    ref_b = torch.cat(O.propagate(adj, Ut, It, 2, include_ego=False)).numpy()
    ref_d = torch.cat(O.propagate(adj, Ut, It, 1)).numpy()
    for r in range(world):
        for k, want_k in (("nccl", ref), ("p2p", ref), ("p2p_b", ref_b), ("p2p_c", ref), ("p2p_d", ref_d), ("p2p_e", ref), ("p2p_g", ref),
                          ("p2p_h", ref), ("p2p_i", ref)):
            err = np.abs(ret[r][k] - want_k).max()
            assert err <= 1e-5 * np.abs(want_k).max(), f"rank {r} {k}: {err}"
        ib, ie = ret[r]["own_items"]
        for lo_, hi_ in ((0, c["n_users"]), (ib, ie)):
            err = np.abs(ret[r]["p2p_f"][lo_:hi_] - ref[lo_:hi_]).max()
            assert err <= 1e-5 * np.abs(ref).max(), f"rank {r} replicate_result=(0,): rows [{lo_},{hi_}) {err}"
        assert ret[r]["need_copies"] < world
        # every GPU holds the same bits, whichever way the rows travelled (TMA bulk copies, SM stores, dense, multicast)
        for k in ("p2p", "p2p_e", "p2p_g", "p2p_h"):
            assert np.array_equal(ret[0]["p2p"], ret[r][k]), f"rank {r} {k} differs from rank 0's TMA-store result"
    print("NVLS multicast available:", ret[0]["has_multicast"])
