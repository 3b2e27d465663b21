"""One-node multi-GPU layer: one process per GPU, ``torch.distributed`` (NCCL over NVLink) as plumbing.

The reference is single-process / single-device; both hot paths shard naturally:

  * scoring  — the item catalogue is split into contiguous ranges, one per GPU (the user table is
    replicated).  Every GPU produces a local top-K for *all* eval users with the fused kernel, the
    (score, id) candidates are exchanged with one all-gather, and each GPU merges the W lists for its
    own slice of the users by (score desc, id asc) — the merge is order independent, so the ids equal
    a single-GPU sweep bit for bit.  Metric partial sums are all-reduced (6 doubles per cut-off).
  * propagation — CSR rows are partitioned by nonzero count; every layer each GPU runs the SpMM on
    its row block and the new embedding rows are all-gathered to rebuild the gather source.  Node ids
    are remapped once to a padded (rank, local row) numbering so the all-gather output *is* the next
    layer's input (no unpack copy).

The compute callables are injectable so the partition / exchange / merge logic is covered on CPU with
the gloo backend (tests/test_dist_gloo.py); the defaults are the CUDA kernels.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

from . import ops
from .evaluator import metrics_from_sums
from .graph import CsrGraph
from .scoring import EvalPlan


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [begin, end) of n things for ``rank`` (first n % world shards get one more)."""
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def partition_rows_by_nnz(rowptr: np.ndarray, world: int) -> List[int]:
    """Row boundaries b[0]=0 <= ... <= b[world]=n_rows with ~nnz/world nonzeros per block."""
    n_rows, nnz = len(rowptr) - 1, int(rowptr[-1])
    bounds = [0]
    for r in range(1, world):
        bounds.append(int(np.searchsorted(rowptr, nnz * r / world, side="left")))
    bounds.append(n_rows)
    return [min(max(b, bounds[i - 1] if i else 0), n_rows) for i, b in enumerate(bounds)]


def _all_gather_stack(t: torch.Tensor, group) -> torch.Tensor:
    world = dist.get_world_size(group)
    out = torch.empty((world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, t.contiguous(), group=group) if t.is_cuda else \
        dist.all_gather(list(out.unbind(0)), t.contiguous(), group=group)
    return out


class ShardedFullRankScorer:
    """Item-sharded full ranking.  ``topk`` returns the merged lists of this rank's user slice."""

    def __init__(self, K: int, precision: int = ops.SCORE_TF32_CHECKED, group=None,
                 local_topk: Optional[Callable] = None, merge: Optional[Callable] = None, metrics: Optional[Callable] = None):
        self.K, self.precision, self.group = int(K), precision, group
        on = dist.is_available() and dist.is_initialized()
        self.rank, self.world = (dist.get_rank(group), dist.get_world_size(group)) if on else (0, 1)
        self.last_n_refined = None
        self._local_topk = local_topk or self._cuda_local_topk
        self._merge = merge or ops.topk_merge
        self._metrics = metrics or (lambda ids, rp, col, Ns: ops.rank_metrics(ids, rp, col, Ns)[0])

    def _cuda_local_topk(self, user_tab, item_tab, item_begin, plan: EvalPlan, item_flags):
        s, i, self.last_n_refined = ops.score_topk(user_tab, item_tab, self.K, user_ids=plan.user_ids, item_id_base=item_begin,
                                 mask_rowptr=plan.mask_rowptr, mask_col=plan.mask_col, item_flags=item_flags,
                                                   flag_exclude=plan.flag_exclude if item_flags is not None else 0,
                                                   precision=self.precision)
        return s, i

    def user_slice(self, n_q: int) -> Tuple[int, int]:
        return shard_range(n_q, self.rank, self.world)

    def topk(self, user_tab, item_shard: torch.Tensor, item_begin: int, plan: EvalPlan, item_flags=None):
        """item_shard = rows [item_begin, item_begin + len) of the item table; plan covers ALL eval users.
        Returns (scores, ids) [n_slice, K] for users user_slice(plan.n_q) — global item ids."""
        s, i = self._local_topk(user_tab, item_shard, item_begin, plan, item_flags)
        if self.world == 1:
            return s, i
        gs, gi = _all_gather_stack(s, self.group), _all_gather_stack(i, self.group)   # [W, n_q, K]
        lo, hi = self.user_slice(plan.n_q)
        return self._merge(gs[:, lo:hi].contiguous(), gi[:, lo:hi].contiguous())

    def metrics(self, ids_slice: torch.Tensor, plan: EvalPlan, Ns: Sequence[int], rounded: bool = True, n_q_total: Optional[int] = None):
        """Hit/Precision/Recall/NDCG over all eval users from per-rank partial sums (one all-reduce).  With ``n_q_total`` the
        plan holds only this rank's user group (see ``GridShardedFullRankScorer.topk``)."""
        n_q = plan.n_q if n_q_total is None else n_q_total
        lo, hi = self.user_slice(n_q) if self.world > 1 else (0, n_q)
        if n_q_total is not None and self.world > 1:          # positions inside the group plan
            g_lo = self.group_slice(n_q)[0] if hasattr(self, "group_slice") else 0
            lo, hi = lo - g_lo, hi - g_lo
        base = plan.gt_rowptr[lo]
        rp = (plan.gt_rowptr[lo:hi + 1] - base).contiguous()
        col = plan.gt_col[int(base):int(plan.gt_rowptr[hi])].contiguous()
        sums = self._metrics(ids_slice, rp, col, list(Ns))
        if self.world > 1:
            dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=self.group)
        return metrics_from_sums(sums.cpu().numpy(), n_q, Ns, rounded)


class UserShardedFullRankScorer(ShardedFullRankScorer):
    """User-sharded full ranking: the item table is replicated (2.56 GB at 10M x 64) and every GPU ranks its slice of the
    eval users against the WHOLE catalogue — no candidate exchange at all, only the metric all-reduce.  Same interface
    as ``ShardedFullRankScorer`` (``topk`` takes the full table with ``item_begin = 0``).  The north-star layout is the
    item-sharded one; this is the zero-communication alternative of SURVEY §8(e): each GPU runs the long sweep (10M
    items per 256-query unit) where the fused kernel is at its best, instead of W times as many short ones."""

    def topk(self, user_tab, item_table: torch.Tensor, item_begin: int, plan: EvalPlan, item_flags=None):
        if item_begin != 0:
            raise ValueError("user-sharded scoring takes the whole item table (item_begin must be 0)")
        lo, hi = self.user_slice(plan.n_q) if self.world > 1 else (0, plan.n_q)
        return self._local_topk(user_tab, item_table, 0, plan.slice(lo, hi) if self.world > 1 else plan, item_flags)


def grid_item_shards(world: int, n_items: int, min_shard_items: int) -> int:
    """Largest divisor S of ``world`` whose shards still hold ``min_shard_items`` items (at least 1)."""
    best = 1
    for s_ in range(1, world + 1):
        if world % s_ == 0 and n_items // s_ >= min_shard_items:
            best = s_
    return best


_SUBGROUPS = {}      # (id of the parent group, S) -> list of sub-groups, created once per process and reused by every scorer


def _user_group_subgroups(group, S: int):
    """The W/S process groups of S consecutive ranks each.  ``dist.new_group`` is collective over the DEFAULT group, so a
    proper sub-group as parent would hang (ranks outside it never make the call): the parent must be WORLD.  The groups are
    cached — constructing a scorer per evaluation must not leak a communicator each time."""
    if group is not None and group is not dist.group.WORLD and dist.get_world_size(group) != dist.get_world_size():
        raise ValueError("GridShardedFullRankScorer with 1 < item_shards < world needs the default (WORLD) process group: "
                         "torch.distributed.new_group is collective over all ranks")
    key = (id(dist.group.WORLD), int(S))      # a re-initialised default group gets fresh sub-groups
    if key not in _SUBGROUPS:
        world = dist.get_world_size()
        _SUBGROUPS[key] = [dist.new_group(ranks=list(range(g_ * S, (g_ + 1) * S))) for g_ in range(world // S)]
    return _SUBGROUPS[key]


class GridShardedFullRankScorer(ShardedFullRankScorer):
    """Item shards x user groups.  The W ranks form W/S user groups of S ranks; inside a group the item catalogue is split
    S ways exactly as in ``ShardedFullRankScorer`` (local top-K, NCCL all-gather of the (score, id) candidates inside the
    group, order-independent merge), and the groups split the eval users.  S = W is the pure item-sharded layout, S = 1 the
    user-sharded one.  Why: the fused sweep pays a fixed price per 256-query unit to warm its thresholds up,
    so its efficiency drops as the shards get short — 0.995 of the tensor peak at 10M items per unit, 0.90 at 2.5M, 0.80 at
    1.25M; on 8 GPUs four 2.5M-item shards x two user groups rank ~12 % more users per second than eight 1.25M-item shards
    while every GPU still holds only a quarter of the catalogue.

    rank r = group r // S, item shard r % S.  ``item_range(n_items)`` is the slice of the item table this rank must hold."""

    def __init__(self, K: int, item_shards: int, precision: int = ops.SCORE_TF32_CHECKED, group=None, **kw):
        super().__init__(K, precision, group, **kw)
        S = int(item_shards)
        if S < 1 or self.world % S != 0:
            raise ValueError(f"item_shards={S} must divide the world size {self.world}")
        self.S, self.n_groups = S, self.world // S
        self.ugroup, self.ishard = self.rank // S, self.rank % S
        self.sub = None
        if self.world > 1 and 1 < S < self.world:
            self.sub = _user_group_subgroups(group, S)[self.ugroup]
        elif S == self.world:
            self.sub = group

    def item_range(self, n_items: int) -> Tuple[int, int]:
        return shard_range(n_items, self.ishard, self.S)

    def group_slice(self, n_q: int) -> Tuple[int, int]:
        return shard_range(n_q, self.ugroup, self.n_groups)

    def user_slice(self, n_q: int) -> Tuple[int, int]:
        g_lo, g_hi = self.group_slice(n_q)
        lo, hi = shard_range(g_hi - g_lo, self.ishard, self.S)
        return g_lo + lo, g_lo + hi

    def topk(self, user_tab, item_shard: torch.Tensor, item_begin: int, plan: EvalPlan, item_flags=None, n_q_total: Optional[int] = None):
        """item_shard = rows ``item_range(n_items)`` of the item table (``item_begin`` = its first id); plan covers ALL
        eval users — or, with ``n_q_total`` given, only this rank's user group ``group_slice(n_q_total)`` (what a host batch
        pipeline copies to this GPU).  Returns (scores, ids) [n_slice, K] for users ``user_slice(n_q)``."""
        g_lo, g_hi = self.group_slice(plan.n_q if n_q_total is None else n_q_total)
        if n_q_total is not None and plan.n_q != g_hi - g_lo:
            raise ValueError(f"group plan has {plan.n_q} users, this rank's group has {g_hi - g_lo}")
        sub_plan = plan if (self.n_groups == 1 or n_q_total is not None) else plan.slice(g_lo, g_hi)
        s, i = self._local_topk(user_tab, item_shard, item_begin, sub_plan, item_flags)
        if self.S == 1:
            return s, i
        gs, gi = _all_gather_stack(s, self.sub), _all_gather_stack(i, self.sub)   # [S, n_group, K]
        lo, hi = shard_range(g_hi - g_lo, self.ishard, self.S)
        return self._merge(gs[:, lo:hi].contiguous(), gi[:, lo:hi].contiguous())


class ShardedItemGenerator:
    """Content -> embedding generators over an item-sharded catalogue.  Tower rows are independent, so
    rank r runs the tower on ITS rows of the content (and backbone) tables only and the generated tile is directly the
    item shard its local scorer sweeps — no collective, tower weights replicated (<= 2.4 MB).  ``scorer`` is a
    ``GridShardedFullRankScorer`` / ``ShardedFullRankScorer`` (the item range is the scorer's).

      * whole-table generators (DropoutNet / Heater item side, model/DropoutNet.py:192-213, model/Heater.py:187-223):
        ``generate(fn, *row_tables)`` = ``fn(*[t[ib:ie] for t in row_tables])``;
      * cold-row overwrite (GAR / ALDI / DeepMusic / MetaEmbedding, model/GAR.py:44-46, model/ALDI.py:89-97):
        ``overwrite_cold(fn, item_shard, content_shard, cold_ids)`` generates only the cold ids that fall in this rank's
        range and scatters them into its shard in place (``fn(content_shard, rows, out)`` — the signature of
        ``towers.gar_generate`` / ``towers.aldi_tower`` with the state bound).
    """

    def __init__(self, scorer, n_items: int):
        self.scorer, self.n_items = scorer, int(n_items)
        self.item_begin, self.item_end = scorer.item_range(n_items) if hasattr(scorer, "item_range") else \
            shard_range(n_items, scorer.rank, scorer.world)

    def rows(self, table):
        """This rank's rows of a full (n_items, *) table; a table that already has shard length is taken as the shard.  A
        ``towers.prepare``d table (``ops.SplitTable``) is sliced half by half."""
        if isinstance(table, ops.SplitTable):
            return ops.SplitTable(self.rows(table.hi), self.rows(table.lo), table.width)
        if table.shape[0] == self.item_end - self.item_begin and table.shape[0] != self.n_items:
            return table
        if table.shape[0] != self.n_items:
            raise ValueError(f"table has {table.shape[0]} rows, expected {self.n_items} (full) or {self.item_end - self.item_begin} (shard)")
        return table[self.item_begin:self.item_end]

    def generate(self, fn: Callable, *row_tables: torch.Tensor) -> torch.Tensor:
        out = fn(*[(lambda r: r if isinstance(r, ops.SplitTable) else r.contiguous())(self.rows(t)) for t in row_tables])
        if out.shape[0] != self.item_end - self.item_begin:
            raise ValueError("the generator must return one row per item of the shard")
        return out

    def local_cold_rows(self, cold_ids) -> torch.Tensor:
        """Cold item ids inside this rank's range, as int32 row numbers of the shard (on the device of ``cold_ids``)."""
        ids = torch.as_tensor(cold_ids)
        sel = ids[(ids >= self.item_begin) & (ids < self.item_end)]
        return (sel - self.item_begin).to(torch.int32).contiguous()

    def overwrite_cold(self, fn: Callable, item_shard: torch.Tensor, content_shard: torch.Tensor, cold_ids) -> torch.Tensor:
        rows = self.local_cold_rows(cold_ids).to(item_shard.device)
        if rows.numel():
            fn(self.rows(content_shard), rows, item_shard)
        return item_shard

    def topk(self, user_tab, item_shard, plan: EvalPlan, item_flags=None):
        """Rank the eval users against the generated catalogue: local sweep over this rank's tile + the scorer's exchange."""
        return self.scorer.topk(user_tab, item_shard, self.item_begin, plan, item_flags)


class RowPartitionedGraph:
    """A square adjacency split by rows over the ranks of ``group`` with padded node numbering.

    Global node n owned by rank r at local offset o gets padded id r * rows_pad + o; the local CSR's
    column ids are remapped accordingly, so a [W * rows_pad, d] all-gather result is directly the
    gather source of the next layer.  ``to_padded`` / ``from_padded`` convert embedding tables.
    """

    def __init__(self, rowptr: np.ndarray, col: np.ndarray, val: Optional[np.ndarray], device, group=None,
                 spmm: Optional[Callable] = None, segments: Optional[Sequence[int]] = None,
                 rank: Optional[int] = None, world: Optional[int] = None, symmetric_alloc: Optional[Callable] = None):
        """``segments`` = row counts of consecutive row classes (for the bipartite adjacency: (user_num,
        item_num)); each class is split by nonzeros on its own and rank r owns part r of every class, which
        balances rows *and* nonzeros (user rows carry ~10x the nonzeros of item rows).  ``rank`` / ``world`` override the
        process group's (one process driving several partitions); ``symmetric_alloc(shape) -> (tensor, handle)`` replaces
        the torch symmetric-memory allocation of the peer-store path (the handle needs ``buffer_ptrs_dev``, ``barrier()``
        and optionally ``multicast_ptr``) — both exist for the CPU emulation in tests/test_dist_gloo.py."""
        self.group = group
        on = dist.is_available() and dist.is_initialized()
        self.rank, self.world = (dist.get_rank(group), dist.get_world_size(group)) if on else (0, 1)
        if rank is not None and world is not None:
            self.rank, self.world = int(rank), int(world)
        self._symmetric_alloc = symmetric_alloc
        rowptr = np.asarray(rowptr)
        self.n = len(rowptr) - 1
        segments = list(segments) if segments is not None else [self.n]
        if sum(segments) != self.n:
            raise ValueError(f"segments {segments} do not add up to {self.n} rows")
        # parts[r] = list of (begin, end) global row ranges owned by rank r, in segment order
        parts = [[] for _ in range(self.world)]
        seg0 = 0
        for n_seg in segments:
            rp = rowptr[seg0:seg0 + n_seg + 1] - rowptr[seg0]
            bounds = partition_rows_by_nnz(rp, self.world)
            for r in range(self.world):
                parts[r].append((seg0 + bounds[r], seg0 + bounds[r + 1]))
            seg0 += n_seg
        self.parts = parts
        self.bounds = [p[0][0] for p in parts] + [self.n] if len(segments) == 1 else None
        sizes = [sum(e - b for b, e in pr) for pr in parts]
        self.rows_pad = int(max(sizes)) if self.n else 0
        self.padded_of = np.empty(self.n, dtype=np.int64)
        self._ranges = []           # (global begin, global end, padded begin): the numbering is piecewise affine
        for r, pr in enumerate(parts):
            off = r * self.rows_pad
            for b, e in pr:
                self.padded_of[b:e] = off + np.arange(e - b)
                if e > b:
                    self._ranges.append((int(b), int(e), int(off)))
                off += e - b
        self.n_local = int(sizes[self.rank])
        # local CSR: this rank's row ranges back to back, then empty padding rows
        mine = parts[self.rank]
        lens = np.concatenate([np.diff(rowptr[b:e + 1]) for b, e in mine]) if mine else np.zeros(0, dtype=np.int64)
        local_rp = np.zeros(self.rows_pad + 1, dtype=np.int64)
        np.cumsum(lens, out=local_rp[1:len(lens) + 1])
        local_rp[len(lens) + 1:] = local_rp[len(lens)]
        col = np.asarray(col)
        local_col = np.concatenate([col[int(rowptr[b]):int(rowptr[e])] for b, e in mine]).astype(np.int64)
        local_col_global = local_col.astype(np.int32)
        local_col = self.padded_of[local_col].astype(np.int32)
        local_val = None
        if val is not None:
            val = np.asarray(val)
            local_val = np.concatenate([val[int(rowptr[b]):int(rowptr[e])] for b, e in mine]).astype(np.float32)
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
        self.local = CsrGraph(t(local_rp), t(local_col), None if local_val is None else t(local_val),
                              self.world * self.rows_pad, row_begin=self.rank * self.rows_pad)
        self._spmm = spmm or (lambda g, X, **kw: g.spmm(X, **kw))
        self._padded_idx = t(self.padded_of)
        # same block with column ids in the reference's node numbering: the first layer of the peer-store path gathers
        # straight from the caller's (N, d) table instead of from a re-ordered copy of it
        self._col_global = t(local_col_global)
        # Sparse all-gather masks (fused peer-store path, W <= 8): bit p of need[r] = "rank p has a nonzero in column r", i.e.
        # p gathers row r in the next layer.  Item rows of a bipartite graph mostly have a handful of nonzeros, hence a
        # handful of readers; only user rows (~100 nonzeros) are read everywhere.
        self._need = self._seg_of_local = None
        self._need_last = {}            # replicate_result -> need mask of the last layer
        self.need_copies = float(self.world)          # mean number of GPUs that read a row (W = dense all-gather)
        if 1 < self.world <= 8 and self.n:
            need = np.zeros(self.n, dtype=np.uint8)
            for r, pr in enumerate(parts):
                seen = np.zeros(self.n, dtype=bool)
                for b, e in pr:
                    seen[col[int(rowptr[b]):int(rowptr[e])]] = True
                need |= seen.astype(np.uint8) << np.uint8(r)
            rows_mine = np.concatenate([np.arange(b, e) for b, e in mine]) if mine else np.zeros(0, dtype=np.int64)
            self._need = t(need[rows_mine] | np.uint8(1 << self.rank))
            self._seg_of_local = np.concatenate([np.full(e - b, k, dtype=np.int64) for k, (b, e) in enumerate(mine)]) if mine else rows_mine
            self.need_copies = float(np.unpackbits(need[:, None], axis=1).sum() / max(self.n, 1))   # mean readers per row

    def to_padded(self, E: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """(N, d) in the reference's node numbering -> (W * rows_pad, d) padded numbering: 2 W contiguous block copies."""
        if out is None:
            out = torch.zeros((self.world * self.rows_pad, E.shape[1]), dtype=E.dtype, device=E.device)
        for b, e, pb in self._ranges:
            out[pb:pb + (e - b)].copy_(E[b:e])
        return out

    def from_padded(self, Ep: torch.Tensor) -> torch.Tensor:
        out = torch.empty((self.n, Ep.shape[1]), dtype=Ep.dtype, device=Ep.device)
        for b, e, pb in self._ranges:
            out[b:e].copy_(Ep[pb:pb + (e - b)])
        return out

    def propagate(self, E0: torch.Tensor, n_layers: int, include_ego: bool = True) -> torch.Tensor:
        """LightGCN-family propagation (model/LightGCN.py:86-96) over the row partition.  E0 is the full
        (N, d) table, replicated; returns the full (N, d) layer mean, identical on every rank."""
        x = self.to_padded(E0)
        r0 = self.rank * self.rows_pad
        count = n_layers + (1 if include_ego else 0)
        acc = torch.empty((self.rows_pad, E0.shape[1]), dtype=E0.dtype, device=E0.device)
        for k in range(1, n_layers + 1):
            last, first = k == n_layers, k == 1
            y = torch.empty_like(acc)
            self._spmm(self.local, x, Y=y, acc=acc, acc_in=(x[r0:r0 + self.rows_pad] if (first and include_ego) else None),
                       acc_beta=(0.0 if (first and not include_ego) else 1.0), acc_div=(float(count) if last else 1.0))
            if not last:
                x = _all_gather_stack(y, self.group).reshape(self.world * self.rows_pad, -1) if self.world > 1 else y
        full = _all_gather_stack(acc, self.group).reshape(self.world * self.rows_pad, -1) if self.world > 1 else acc
        return self.from_padded(full)

    # ---- fused SpMM + all-gather over NVLink peer memory --------------------------------------------------
    def enable_p2p(self, d: int) -> None:
        """Allocate three symmetric (peer-mapped) padded tables [W*rows_pad, d]: two ping-pong gather sources and
        one for the finished layer mean.  Uses torch symmetric memory only as the allocator / pointer exchange /
        inter-GPU barrier; the data movement is done by the SpMM kernel's own epilogue stores."""
        dev = self.local.rowptr.device
        self._p2p_d, self._xbuf, self._xhdl = d, [], []
        shape = (self.world * self.rows_pad, d)
        for _ in range(3):
            if self._symmetric_alloc is not None:
                t, h = self._symmetric_alloc(shape)
            else:
                import torch.distributed._symmetric_memory as symm_mem
                group = self.group if self.group is not None else dist.group.WORLD
                t = symm_mem.empty(shape, dtype=torch.float32, device=dev)
                h = symm_mem.rendezvous(t, group=group)
            self._xhdl.append(h)
            self._xbuf.append(t)
        self.has_multicast = all(int(getattr(h, "multicast_ptr", 0) or 0) != 0 for h in self._xhdl)

    def propagate_p2p(self, E0: torch.Tensor, n_layers: int, include_ego: bool = True, padded_io: bool = False,
                      copy: bool = True, sparse: bool = True, replicate_result: Optional[Sequence[int]] = None,
                      multicast: bool = False, layer_events: Optional[list] = None) -> torch.Tensor:
        """Same result as ``propagate``, but every finished row is stored by the SpMM epilogue straight into the
        gather table of the GPUs that read it (``cr_spmm_csr_bcast_f32``): the per-layer all-gather overlaps the SpMM
        instead of following it, and with ``sparse`` it only moves a row to the GPUs whose row block has a nonzero in
        that column.  The last layer's epilogue scatters the finished layer mean into every GPU's result table *in the
        reference's node numbering* (two destination ranges per rank: its user rows and its item rows), so no
        re-ordering pass follows.  ``replicate_result`` = indices of the row classes (``segments``) whose result rows
        every GPU receives — default all of them; ``(0,)`` replicates the user rows only and leaves each item row
        with its owner, which is what item-sharded scoring consumes (rows of other owners are then undefined).
        With ``multicast`` (and an NVLS-capable node) rows wanted by every GPU leave as ONE ``multimem.st`` on the
        table's multicast address and are replicated by the NVSwitch instead of being stored W times — off by default:
        measured slower than the unicast stores at 16 bytes per lane (C4, 4 GPUs: 12.2 vs 11.0 ms per step).
        With ``padded_io`` E0 and the result are in the padded numbering instead.  With ``copy=False`` the result is
        a view of the peer-mapped result table, valid until the next call."""
        if getattr(self, "_p2p_d", None) != E0.shape[1]:
            self.enable_p2p(E0.shape[1])
        W, r0, nl = self.world, self.rank * self.rows_pad, self.n_local
        src, hdl = self._xbuf, self._xhdl
        hdl[2].barrier()                     # nobody still reads the buffers of a previous call
        count = n_layers + (1 if include_ego else 0)
        acc = torch.empty((self.rows_pad, E0.shape[1]), dtype=E0.dtype, device=E0.device)
        mine = [(b, e) for b, e in self.parts[self.rank]]
        if padded_io:
            src[0].copy_(E0)
            x0, col0 = src[0], self.local.col
            if include_ego:
                acc[:nl].copy_(E0[r0:r0 + nl])
        else:                                # layer 1 reads the caller's table in place (column ids in its numbering)
            x0, col0 = E0.contiguous(), self._col_global
            o = 0
            for b, e in mine:
                if include_ego:
                    acc[o:o + (e - b)].copy_(E0[b:e])
                o += e - b
        plan = self.local.plan(E0.shape[1]) if self.local.rowptr.is_cuda else None      # (CPU: emulated kernel, no plan)
        rowptr = self.local.rowptr[:nl + 1]  # the padding rows are empty and nobody gathers them: not computed, not sent
        scatter = (not padded_io) and len(mine) <= 2
        need = self._need if sparse else None
        need_last = None
        if replicate_result is not None and self._need is not None:
            key = tuple(sorted(replicate_result))
            cache = self._need_last
            if key not in cache:
                m = np.full(nl, 1 << self.rank, dtype=np.uint8)
                m[np.isin(self._seg_of_local, key)] = (1 << W) - 1
                cache[key] = torch.from_numpy(m).to(E0.device)
            need_last = cache[key]
        for k in range(1, n_layers + 1):
            last, first = k == n_layers, k == 1
            x = x0 if first else src[(k - 1) % 2]
            out_h = hdl[2] if last else hdl[k % 2]
            off, split, off_hi = r0, None, 0
            if last and scatter:             # local rows [0, n0) are global rows [b0, e0); the rest are [b1, e1)
                n0 = mine[0][1] - mine[0][0]
                off, split = mine[0][0], n0
                off_hi = (mine[1][0] - n0) if len(mine) == 2 else 0
            if layer_events is not None:     # (probe) CUDA events around each layer's kernels and its barrier
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                layer_events.append(ev)
                ev[0].record()
            ops.spmm_bcast(rowptr, col0 if first else self.local.col, self.local.val, x, out_h.buffer_ptrs_dev, W, off, acc=acc,
                           acc_beta=(0.0 if (first and not include_ego) else 1.0), acc_div=(float(count) if last else 1.0),
                           plan=plan, bcast_acc=last, peer_row_split=split, peer_row_offset_hi=off_hi,
                           peer_need=(need_last if last else need),
                           multicast_ptr=(int(getattr(out_h, "multicast_ptr", 0) or 0) if multicast else 0))
            if layer_events is not None:
                layer_events[-1][1].record()
            out_h.barrier()                  # every GPU's rows have landed everywhere
            if layer_events is not None:
                layer_events[-1][2].record()
        if padded_io:
            res = src[2]
        else:
            res = src[2][:self.n] if scatter else self.from_padded(src[2])
        return res.clone() if (copy and res.data_ptr() == src[2].data_ptr()) else res
