"""world_size-2 gloo tests (CPU) of the multi-GPU host logic in coldrec_b200/dist.py: item sharding +
candidate all-gather + per-slice merge + metric all-reduce, and row-partitioned propagation with padded
node numbering.  The CUDA kernels are replaced by oracle-backed callables (the injection points exist for
exactly this); results must equal the single-process oracle."""
import os
import socket

import numpy as np
import pytest
import scipy.sparse as sp
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import coldrec_oracle as O


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _case():
    rng = np.random.default_rng(123)
    n_users, n_items, n_q, K = 120, 900, 61, 20
    U = (rng.standard_normal((n_users, 64)) * 0.2).astype(np.float32)
    I = (rng.standard_normal((n_items, 64)) * 0.2).astype(np.float32)
    I[rng.choice(n_items, 40, replace=False)] = I[rng.choice(n_items, 40, replace=False)]      # exact ties
    uids = rng.choice(n_users, n_q, replace=False).astype(np.int32)
    rows = [np.sort(rng.choice(n_items, int(rng.integers(0, 30)), replace=False)) for _ in range(n_q)]
    rowptr = np.zeros(n_q + 1, dtype=np.int64); np.cumsum([len(r) for r in rows], out=rowptr[1:])
    col = np.concatenate(rows).astype(np.int32)
    gt = [np.sort(rng.choice(n_items, int(rng.integers(0, 8)), replace=False)) for _ in range(n_q)]
    gt_rowptr = np.zeros(n_q + 1, dtype=np.int64); np.cumsum([len(r) for r in gt], out=gt_rowptr[1:])
    gt_col = np.concatenate(gt).astype(np.int32)
    eu, ei = rng.integers(0, n_users, 3000), rng.integers(0, n_items, 3000)
    return dict(U=U, I=I, uids=uids, rowptr=rowptr, col=col, gt_rowptr=gt_rowptr, gt_col=gt_col, K=K, eu=eu, ei=ei,
                n_users=n_users, n_items=n_items)


def _sorted_topk(scores, ids, K):
    """(score desc, id asc) top-K of candidate rows, padding ids < 0 last — the merge kernel's order."""
    out_s = np.full((scores.shape[0], K), -np.inf, dtype=np.float32)
    out_i = np.full((scores.shape[0], K), -1, dtype=np.int32)
    for j in range(scores.shape[0]):
        valid = ids[j] >= 0
        order = np.lexsort((ids[j][valid], -scores[j][valid].astype(np.float64)))[:K]
        out_s[j, :len(order)] = scores[j][valid][order]
        out_i[j, :len(order)] = ids[j][valid][order]
    return out_s, out_i


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from coldrec_b200.dist import RowPartitionedGraph, ShardedFullRankScorer, UserShardedFullRankScorer, shard_range
        from coldrec_b200.scoring import EvalPlan
        c = _case()
        K = c["K"]
        Ut, It = torch.from_numpy(c["U"]), torch.from_numpy(c["I"])

        def local_topk(user_tab, item_tab, item_begin, plan, item_flags):
            # exact (score desc, id asc) local list from oracle scores of this shard
            n_loc = item_tab.shape[0]
            s = (user_tab[plan.user_ids.long()] @ item_tab.T).numpy().copy()
            prp, pcol = plan.mask_rowptr.numpy(), plan.mask_col.numpy()
            for j in range(plan.n_q):
                m = pcol[prp[j]:prp[j + 1]] - item_begin
                s[j, m[(m >= 0) & (m < n_loc)]] = O.MASK_SENTINEL
            ids = np.broadcast_to(np.arange(item_begin, item_begin + n_loc, dtype=np.int32), s.shape)
            ts, ti = _sorted_topk(s, ids, K)
            return torch.from_numpy(ts), torch.from_numpy(ti)

        def merge(gs, gi):
            W, n, k = gs.shape
            ts, ti = _sorted_topk(gs.permute(1, 0, 2).reshape(n, W * k).numpy(), gi.permute(1, 0, 2).reshape(n, W * k).numpy(), k)
            return torch.from_numpy(ts), torch.from_numpy(ti)

        def metrics(ids, rp, col, Ns):
            n = ids.shape[0]
            sums = np.zeros((len(Ns), 6))
            for a, N in enumerate(Ns):
                for j in range(n):
                    g = set(col[rp[j]:rp[j + 1]].tolist())
                    row = ids[j, :N].tolist()
                    hits = len(g.intersection(row))
                    dcg = sum(1.0 / np.log2(k + 2) for k, it in enumerate(row) if it in g)
                    idcg = sum(1.0 / np.log2(k + 2) for k in range(min(len(g), N)))
                    sums[a] += [hits, len(g), hits / len(g) if g else 0, 1 if g else 0, dcg / idcg if idcg else 0, 1 if idcg else 0]
            return torch.from_numpy(sums)

        plan = EvalPlan.from_arrays(torch.from_numpy(c["uids"]), torch.from_numpy(c["rowptr"]), torch.from_numpy(c["col"]),
                                    torch.from_numpy(c["gt_rowptr"]), torch.from_numpy(c["gt_col"]))
        sc = ShardedFullRankScorer(K, group=None, local_topk=local_topk, merge=merge, metrics=metrics)
        b, e = shard_range(c["n_items"], rank, world)
        s, i = sc.topk(Ut, It[b:e], b, plan)
        perf = sc.metrics(i, plan, [10, 20], rounded=False)
        lo, hi = sc.user_slice(plan.n_q)
        usc = UserShardedFullRankScorer(K, group=None, local_topk=local_topk, merge=merge, metrics=metrics)
        us, ui = usc.topk(Ut, It, 0, plan)
        uperf = usc.metrics(ui, plan, [10, 20], rounded=False)

        adj = O.normalize_graph_mat(O.bipartite_adjacency(c["eu"], c["ei"], c["n_users"], c["n_items"])).tocsr()
        adj.sort_indices()

        def cpu_spmm(g, X, Y=None, acc=None, acc_in=None, acc_beta=1.0, acc_div=1.0):
            A = sp.csr_matrix((g.val.numpy(), g.col.numpy(), g.rowptr.numpy()), shape=(g.n_rows, g.n_cols))
            y = torch.from_numpy(A @ X.numpy())
            if Y is not None:
                Y.copy_(y)
            if acc is not None:
                src = acc_in if acc_in is not None else acc
                acc.copy_(((acc_beta * src + y) if acc_beta else y) / acc_div)

        G = RowPartitionedGraph(adj.indptr.astype(np.int64), adj.indices.astype(np.int64), adj.data.astype(np.float32), "cpu",
                                spmm=cpu_spmm)
        G2 = RowPartitionedGraph(adj.indptr.astype(np.int64), adj.indices.astype(np.int64), adj.data.astype(np.float32), "cpu",
                                 spmm=cpu_spmm, segments=(c["n_users"], c["n_items"]))
        E0 = torch.cat([Ut, It])
        out = {}
        for ego in (True, False):
            out[ego] = G.propagate(E0, 3, include_ego=ego).numpy()
        out["seg"] = G2.propagate(E0, 3).numpy()
        local = dict(rowptr=G2.local.rowptr.numpy(), col=G2.local.col.numpy(), val=G2.local.val.numpy(), need=G2._need.numpy(),
                     n_local=G2.n_local, padded_of=G2.padded_of)
        ret[rank] = dict(s=s.numpy(), i=i.numpy(), lo=lo, hi=hi, perf=perf, prop=out, bounds=G.bounds, parts=G2.parts,
                         rows_pad=G2.rows_pad, us=us.numpy(), ui=ui.numpy(), uperf=uperf, local=local)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_scoring_and_partitioned_propagation_world2():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    c = _case()
    K = c["K"]
    Ut, It = torch.from_numpy(c["U"]), torch.from_numpy(c["I"])
    # single-process truth with the same deterministic order
    s_full = (Ut[torch.from_numpy(c["uids"]).long()] @ It.T).numpy().copy()
    for j in range(len(c["uids"])):
        s_full[j, c["col"][c["rowptr"][j]:c["rowptr"][j + 1]]] = O.MASK_SENTINEL
    ids = np.broadcast_to(np.arange(c["n_items"], dtype=np.int32), s_full.shape)
    ts, ti = _sorted_topk(s_full, ids, K)
    covered = 0
    for r in range(world):
        lo, hi = ret[r]["lo"], ret[r]["hi"]
        assert np.array_equal(ret[r]["i"], ti[lo:hi]), "sharded ids must equal the single sweep bit for bit"
        assert np.allclose(ret[r]["s"], ts[lo:hi], atol=1e-6)
        assert np.array_equal(ret[r]["ui"], ti[lo:hi]), "user-sharded ids must equal the single sweep as well"
        assert np.allclose(ret[r]["us"], ts[lo:hi], atol=1e-6)
        covered += hi - lo
    assert covered == len(c["uids"])
    want = O.metrics_from_topk(ti.astype(np.int64), c["gt_rowptr"], c["gt_col"].astype(np.int64), [10, 20])
    for r in range(world):
        assert np.allclose(ret[r]["perf"], want, atol=1e-9), "all-reduced metrics must equal the global ones on every rank"
        assert np.allclose(ret[r]["uperf"], want, atol=1e-9)
    adj = O.normalize_graph_mat(O.bipartite_adjacency(c["eu"], c["ei"], c["n_users"], c["n_items"]))
    for ego in (True, False):
        ru, ri = O.propagate(adj, Ut, It, 3, include_ego=ego)
        ref = torch.cat([ru, ri]).numpy()
        for r in range(world):
            assert np.abs(ret[r]["prop"][ego] - ref).max() <= 1e-5 * np.abs(ref).max()
    assert ret[0]["bounds"] == ret[1]["bounds"] and ret[0]["bounds"][0] == 0
    ru, ri = O.propagate(adj, Ut, It, 3)
    ref = torch.cat([ru, ri]).numpy()
    for r in range(world):     # two-segment partition: every rank owns a slice of the users and a slice of the items
        assert np.abs(ret[r]["prop"]["seg"] - ref).max() <= 1e-5 * np.abs(ref).max()
        (ub, ue), (ib, ie) = ret[r]["parts"][r]
        assert 0 <= ub < ue <= c["n_users"] <= ib < ie <= c["n_users"] + c["n_items"]
    assert ret[0]["rows_pad"] <= 0.6 * (c["n_users"] + c["n_items"])
    _check_sparse_allgather(ret, world, torch.cat([Ut, It]).numpy(), ref)


def _check_sparse_allgather(ret, world, E0, ref):
    """The fused peer-store path only sends a row to the ranks whose need bit is set (cr_spmm_csr_bcast_f32, peer_need).
    Emulated here with the workers' own local CSRs and masks: every table row a rank did not receive is NaN, so a row that
    is read without having been sent poisons the result."""
    rows_pad, N = ret[0]["rows_pad"], E0.shape[0]
    loc = [ret[r]["local"] for r in range(world)]
    padded_of = loc[0]["padded_of"]
    A = [sp.csr_matrix((l["val"], l["col"], l["rowptr"]), shape=(rows_pad, world * rows_pad)) for l in loc]
    # the masks are exact: bit p of need[row] <=> rank p's block has a nonzero in that (padded) column, or p owns the row
    for r in range(world):
        for p_ in range(world):
            cols = np.zeros(world * rows_pad, dtype=bool)
            cols[loc[p_]["col"]] = True
            want = cols[r * rows_pad:r * rows_pad + loc[r]["n_local"]] | (p_ == r)
            assert np.array_equal(((loc[r]["need"] >> p_) & 1).astype(bool), want)
    full = np.zeros((world * rows_pad, E0.shape[1]), dtype=np.float64)
    full[padded_of] = E0
    tables = [full.copy() for _ in range(world)]
    acc = [full[r * rows_pad:(r + 1) * rows_pad].copy() for r in range(world)]
    for _ in range(3):
        new = [np.full_like(full, np.nan) for _ in range(world)]
        for r in range(world):
            with np.errstate(invalid="ignore"):
                # scipy multiplies explicit nonzeros only, like the kernel: NaN rows outside the pattern are never touched
                y = A[r] @ np.where(np.isnan(tables[r]), 0.0, tables[r])
                poisoned = (abs(A[r]).astype(bool).astype(np.float64) @ np.isnan(tables[r]).any(1).astype(np.float64)) > 0
            assert not poisoned.any(), "a row was gathered that had not been sent to this rank"
            acc[r] += y
            n_l = loc[r]["n_local"]
            for p_ in range(world):
                sel = np.nonzero((loc[r]["need"] >> p_) & 1)[0]
                new[p_][r * rows_pad + sel] = y[:n_l][sel]
        tables = new
    got = np.concatenate([a / 4.0 for a in acc])[padded_of]
    assert np.abs(got - ref).max() <= 1e-5 * np.abs(ref).max()
    sent = sum(int(np.unpackbits(l["need"][:, None], axis=1).sum()) for l in loc)
    assert sent < world * N, "the masks should save some copies on this graph"


def _scores(Uq, It):
    """fp32 scores that do not depend on how the matrices are blocked: the CUDA scorer sums k = 0..63 in a fixed order whatever
    the shard shape, a CPU sgemm does not (its last bit moves with the matrix shape, which reorders near-ties between a shard
    and the single sweep) — so accumulate in fp64 and round once."""
    return (Uq.double() @ It.double().T).float().numpy().copy()


def _cpu_scoring_callables(K):
    """Oracle-backed stand-ins for the three CUDA entry points of the sharded scorer (same order as the kernels)."""
    def local_topk(user_tab, item_tab, item_begin, plan, item_flags):
        n_loc = item_tab.shape[0]
        s = _scores(user_tab[plan.user_ids.long()], item_tab)
        prp, pcol = plan.mask_rowptr.numpy(), plan.mask_col.numpy()
        for j in range(plan.n_q):
            m = pcol[prp[j]:prp[j + 1]] - item_begin
            s[j, m[(m >= 0) & (m < n_loc)]] = O.MASK_SENTINEL
        ids = np.broadcast_to(np.arange(item_begin, item_begin + n_loc, dtype=np.int32), s.shape)
        ts, ti = _sorted_topk(s, ids, K)
        return torch.from_numpy(ts), torch.from_numpy(ti)

    def merge(gs, gi):
        W, n, k = gs.shape
        ts, ti = _sorted_topk(gs.permute(1, 0, 2).reshape(n, W * k).numpy(), gi.permute(1, 0, 2).reshape(n, W * k).numpy(), k)
        return torch.from_numpy(ts), torch.from_numpy(ti)

    def metrics(ids, rp, col, Ns):
        sums = np.zeros((len(Ns), 6))
        for a, N in enumerate(Ns):
            for j in range(ids.shape[0]):
                g = set(col[rp[j]:rp[j + 1]].tolist())
                row = ids[j, :N].tolist()
                hits = len(g.intersection(row))
                dcg = sum(1.0 / np.log2(k + 2) for k, it in enumerate(row) if it in g)
                idcg = sum(1.0 / np.log2(k + 2) for k in range(min(len(g), N)))
                sums[a] += [hits, len(g), hits / len(g) if g else 0, 1 if g else 0, dcg / idcg if idcg else 0, 1 if idcg else 0]
        return torch.from_numpy(sums)
    return dict(local_topk=local_topk, merge=merge, metrics=metrics)


def _grid_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from coldrec_b200.dist import GridShardedFullRankScorer
        from coldrec_b200.scoring import EvalPlan
        c = _case()
        Ut, It = torch.from_numpy(c["U"]), torch.from_numpy(c["I"])
        plan = EvalPlan.from_arrays(torch.from_numpy(c["uids"]), torch.from_numpy(c["rowptr"]), torch.from_numpy(c["col"]),
                                    torch.from_numpy(c["gt_rowptr"]), torch.from_numpy(c["gt_col"]))
        out = {}
        for S in [s_ for s_ in (1, 2, 4, 8) if world % s_ == 0]:     # user-sharded ... grids ... pure item-sharded
            sc = GridShardedFullRankScorer(c["K"], S, **_cpu_scoring_callables(c["K"]))
            b, e = sc.item_range(c["n_items"])
            s, i = sc.topk(Ut, It[b:e], b, plan)
            lo, hi = sc.user_slice(plan.n_q)
            # a host pipeline hands each rank only its user group's rows of the plan (HostBatchEvaluator): same lists, same metrics
            g_lo, g_hi = sc.group_slice(plan.n_q)
            s_g, i_g = sc.topk(Ut, It[b:e], b, plan.slice(g_lo, g_hi), n_q_total=plan.n_q)
            assert torch.equal(i_g, i) and torch.equal(s_g, s)
            perf_g = sc.metrics(i_g, plan.slice(g_lo, g_hi), [10, 20], rounded=False, n_q_total=plan.n_q)
            out[S] = dict(s=s.numpy(), i=i.numpy(), lo=lo, hi=hi, perf=sc.metrics(i, plan, [10, 20], rounded=False), items=(b, e), perf_g=perf_g)
        ret[rank] = out
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("world", [4, 8])
def test_grid_sharded_scoring(world):
    """Item shards x user groups on 4 and 8 ranks (8 ranks, S=4 is the bench's 8-GPU layout): every user is ranked exactly once,
    lists equal the single sweep bit for bit, all-reduced metrics equal the global ones."""
    from coldrec_b200.dist import grid_item_shards
    ret = mp.Manager().dict()
    mp.spawn(_grid_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    c = _case()
    K = c["K"]
    Ut, It = torch.from_numpy(c["U"]), torch.from_numpy(c["I"])
    s_full = _scores(Ut[torch.from_numpy(c["uids"]).long()], It)
    for j in range(len(c["uids"])):
        s_full[j, c["col"][c["rowptr"][j]:c["rowptr"][j + 1]]] = O.MASK_SENTINEL
    ts, ti = _sorted_topk(s_full, np.broadcast_to(np.arange(c["n_items"], dtype=np.int32), s_full.shape), K)
    want = O.metrics_from_topk(ti.astype(np.int64), c["gt_rowptr"], c["gt_col"].astype(np.int64), [10, 20])
    for S in [s_ for s_ in (1, 2, 4, 8) if world % s_ == 0]:
        seen = np.zeros(len(c["uids"]), dtype=int)
        for r in range(world):
            o = ret[r][S]
            assert np.array_equal(o["i"], ti[o["lo"]:o["hi"]]), f"S={S} rank {r}: ids differ from the single sweep"
            assert np.allclose(o["s"], ts[o["lo"]:o["hi"]], atol=1e-6)
            assert np.allclose(o["perf"], want, atol=1e-9) and np.allclose(o["perf_g"], want, atol=1e-9)
            seen[o["lo"]:o["hi"]] += 1
            assert o["items"][1] - o["items"][0] in (c["n_items"] // S, c["n_items"] // S + 1)
        assert (seen == 1).all(), f"S={S}: the user slices must tile the eval users"
    assert [grid_item_shards(w, 10_000_000, 2_500_000) for w in (1, 2, 4, 8)] == [1, 2, 4, 4]
    assert grid_item_shards(8, 10_000_000, 1) == 8 and grid_item_shards(8, 100, 1000) == 1 and grid_item_shards(6, 900, 250) == 3


class _FakeSymmetric:
    """Stand-in for torch symmetric memory in ONE process: buffer k of every rank is registered under the same key, a
    barrier is a threading.Barrier.  `buffer_ptrs_dev` is the key the emulated kernel uses to find the peers' tables."""

    def __init__(self, world):
        import threading
        self.world, self.tables, self.bar = world, {}, threading.Barrier(world)
        self.count = [0] * world

    def allocator(self, rank):
        def alloc(shape):
            key = self.count[rank]
            self.count[rank] += 1
            t = torch.full(shape, float("nan"), dtype=torch.float32)
            self.tables.setdefault(key, {})[rank] = t
            outer = self

            class H:
                buffer_ptrs_dev, multicast_ptr = key, 0

                def barrier(self_inner):
                    outer.bar.wait()
            return t, H()
        return alloc


def _emulated_spmm_bcast(fake):
    """cr_spmm_csr_bcast_f32 in numpy, store_epilogue semantics included (coldrec_b200/csrc/spmm.cu): y = A_local X;
    acc = (beta * acc + y) / div; row r goes to destination row off + r (r < split) or off_hi + r, on the peers whose need bit
    is set (all peers without a mask); with bcast_acc the peers receive acc instead of y."""
    def spmm_bcast(rowptr, col, val, X, peer_tables_dev, n_peers, peer_row_offset, acc=None, acc_in=None, acc_beta=1.0, acc_div=1.0,
                   plan=None, bcast_acc=False, peer_row_split=None, peer_row_offset_hi=0, peer_need=None, multicast_ptr=0):
        n = rowptr.numel() - 1
        nnz = int(rowptr[-1])
        A = sp.csr_matrix((val.numpy()[:nnz], col.numpy()[:nnz], rowptr.numpy()), shape=(n, X.shape[0]))
        Xn = X.numpy()
        assert not np.isnan(Xn[np.unique(col.numpy()[:nnz])]).any(), "gathered a row that was never delivered to this rank"
        y = torch.from_numpy((A @ np.nan_to_num(Xn)).astype(np.float32))
        out = y
        if acc is not None:
            src = acc_in if acc_in is not None else acc
            res = ((acc_beta * src[:n] + y) if acc_beta else y) / acc_div
            acc[:n] = res
            if bcast_acc:
                out = res
        split = n if peer_row_split is None else peer_row_split
        r = np.arange(n)
        dest = torch.from_numpy(np.where(r < split, peer_row_offset + r, peer_row_offset_hi + r))
        need = peer_need.numpy()[:n] if peer_need is not None else np.full(n, (1 << n_peers) - 1)
        for p_ in range(n_peers):
            sel = torch.from_numpy(np.nonzero((need >> p_) & 1)[0])
            fake.tables[peer_tables_dev][p_][dest[sel]] = out[sel]
        return acc
    return spmm_bcast


@pytest.mark.parametrize("world", [2, 4, 8])
def test_peer_store_propagation_host_logic_emulated(world, monkeypatch):
    """RowPartitionedGraph.propagate_p2p on CPU with the kernel and symmetric memory emulated (one thread per rank): need masks,
    first layer read in place, ping-pong tables, last-layer scatter into the reference numbering (two ranges per rank),
    users-only result replication — for 2, 4 and 8 ranks (the GPU test covers 2)."""
    import threading
    from coldrec_b200 import dist as crd
    c = _case()
    adj = O.normalize_graph_mat(O.bipartite_adjacency(c["eu"], c["ei"], c["n_users"], c["n_items"])).tocsr()
    adj.sort_indices()
    Ut, It = torch.from_numpy(c["U"]), torch.from_numpy(c["I"])
    E0 = torch.cat([Ut, It])
    fake = _FakeSymmetric(world)
    monkeypatch.setattr(crd.ops, "spmm_bcast", _emulated_spmm_bcast(fake))
    graphs = [crd.RowPartitionedGraph(adj.indptr.astype(np.int64), adj.indices.astype(np.int64), adj.data.astype(np.float32), "cpu",
                                      segments=(c["n_users"], c["n_items"]), rank=r, world=world, symmetric_alloc=fake.allocator(r))
              for r in range(world)]
    results, errors = {}, []

    def run(r):
        try:
            G = graphs[r]
            results[r] = dict(full=G.propagate_p2p(E0, 3).numpy().copy(),
                              noego=G.propagate_p2p(E0, 2, include_ego=False).numpy().copy(),
                              dense=G.propagate_p2p(E0, 3, sparse=False).numpy().copy(),
                              one=G.propagate_p2p(E0, 1).numpy().copy(),
                              users=G.propagate_p2p(E0, 3, replicate_result=(0,)).numpy().copy(),
                              padded=G.from_padded(G.propagate_p2p(G.to_padded(E0), 3, padded_io=True)).numpy().copy())
        except Exception as ex:          # a dead thread would leave the others waiting at the barrier
            errors.append((r, repr(ex)))
            fake.bar.abort()
    threads = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    [t_.start() for t_ in threads]
    [t_.join(timeout=120) for t_ in threads]
    assert not errors, errors
    ref = {L: torch.cat(O.propagate(adj, Ut, It, L)).numpy() for L in (1, 3)}
    ref_noego = torch.cat(O.propagate(adj, Ut, It, 2, include_ego=False)).numpy()
    tol = lambda want: 1e-5 * np.abs(want).max()
    for r in range(world):
        for key, want in (("full", ref[3]), ("dense", ref[3]), ("padded", ref[3]), ("one", ref[1]), ("noego", ref_noego)):
            got = results[r][key]
            assert not np.isnan(got).any(), f"rank {r} {key}: rows missing from the result"
            assert np.abs(got - want).max() <= tol(want), f"rank {r} {key}"
        (ub, ue), (ib, ie) = graphs[r].parts[r]
        got = results[r]["users"]
        for lo_, hi_ in ((0, c["n_users"]), (ib, ie)):          # user rows everywhere, item rows with their owner
            assert np.abs(got[lo_:hi_] - ref[3][lo_:hi_]).max() <= tol(ref[3]), f"rank {r} users-only rows [{lo_},{hi_})"
    assert graphs[0].need_copies < world


def test_partition_helpers():
    from coldrec_b200.dist import partition_rows_by_nnz, shard_range
    assert [shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert shard_range(3, 3, 4) == (3, 3)
    rowptr = np.array([0, 100, 100, 101, 150, 200, 200, 400])
    b = partition_rows_by_nnz(rowptr, 4)
    assert b[0] == 0 and b[-1] == 7 and all(x <= y for x, y in zip(b, b[1:]))
    nnz = [rowptr[b[i + 1]] - rowptr[b[i]] for i in range(4)]
    assert max(nnz) <= 200


# ---------------------------------------------------------------------------------------------- sharded generators
def _gen_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from coldrec_b200.dist import GridShardedFullRankScorer, ShardedItemGenerator
        from coldrec_b200.scoring import EvalPlan
        c = _case()
        Ut, It = torch.from_numpy(c["U"]), torch.from_numpy(c["I"])
        plan = EvalPlan.from_arrays(torch.from_numpy(c["uids"]), torch.from_numpy(c["rowptr"]), torch.from_numpy(c["col"]),
                                    torch.from_numpy(c["gt_rowptr"]), torch.from_numpy(c["gt_col"]))
        content, W1, cold = _gen_inputs(c)
        out = {}
        for S in [s_ for s_ in (1, 2, 4) if world % s_ == 0]:
            sc = GridShardedFullRankScorer(c["K"], S, **_cpu_scoring_callables(c["K"]))
            gen = ShardedItemGenerator(sc, c["n_items"])
            # whole-table generator (DropoutNet / Heater style): [V | content] rows of this rank only
            shard = gen.generate(lambda V, C: _tower(torch.cat([V, C], 1), W1), It, content)
            s, i = gen.topk(Ut, shard, plan)
            # cold-row overwrite (GAR / ALDI style) into this rank's slice of the backbone table
            base = gen.rows(It).clone()
            def scatter(cont, rows, out_):
                out_[rows.long()] = _tower(torch.cat([out_[rows.long()], cont[rows.long()]], 1), W1)
            gen.overwrite_cold(scatter, base, content, cold)
            s2, i2 = gen.topk(Ut, base, plan)
            lo, hi = sc.user_slice(plan.n_q)
            out[S] = dict(i=i.numpy(), s=s.numpy(), i2=i2.numpy(), lo=lo, hi=hi, rows=(gen.item_begin, gen.item_end),
                          n_cold_local=int(gen.local_cold_rows(cold).numel()))
        ret[rank] = out
    finally:
        dist.destroy_process_group()


def _tower(x, W1):
    return torch.tanh(x.double() @ W1.double().T).float()


def _gen_inputs(c):
    g = torch.Generator().manual_seed(12)
    content = torch.randn(c["n_items"], 24, generator=g)
    W1 = torch.randn(64, 64 + 24, generator=g) * 0.2
    cold = torch.sort(torch.randperm(c["n_items"], generator=g)[:c["n_items"] // 5]).values
    return content, W1, cold


@pytest.mark.timeout(600)
@pytest.mark.parametrize("world", [2, 4])
def test_sharded_generators_feed_the_local_scorer(world):
    """each rank generates only its item range (whole-table towers and the cold-row overwrite); the ranked
    lists equal generating the whole catalogue on one device and sweeping it once."""
    ret = mp.Manager().dict()
    mp.spawn(_gen_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    c = _case()
    K = c["K"]
    Ut, It = torch.from_numpy(c["U"]), torch.from_numpy(c["I"])
    content, W1, cold = _gen_inputs(c)
    full_a = _tower(torch.cat([It, content], 1), W1)
    full_b = It.clone()
    full_b[cold] = _tower(torch.cat([It[cold], content[cold]], 1), W1)
    want = []
    for tab in (full_a, full_b):
        sc = _scores(Ut[torch.from_numpy(c["uids"]).long()], tab)
        for j in range(len(c["uids"])):
            sc[j, c["col"][c["rowptr"][j]:c["rowptr"][j + 1]]] = O.MASK_SENTINEL
        want.append(_sorted_topk(sc, np.broadcast_to(np.arange(c["n_items"], dtype=np.int32), sc.shape), K)[1])
    for S in [s_ for s_ in (1, 2, 4) if world % s_ == 0]:
        n_cold = {}
        for r in range(world):
            o = ret[r][S]
            assert np.array_equal(o["i"], want[0][o["lo"]:o["hi"]]), f"S={S} rank {r}: whole-table generator"
            assert np.array_equal(o["i2"], want[1][o["lo"]:o["hi"]]), f"S={S} rank {r}: cold-row overwrite"
            n_cold[o["rows"]] = o["n_cold_local"]
        assert sum(n_cold.values()) == len(cold), "every cold item is generated by exactly one item shard"
