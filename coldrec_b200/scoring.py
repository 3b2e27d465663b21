"""Full-catalogue scoring: eval plans and the fused scorer front end.

``EvalPlan`` replaces ``BaseColdStartTrainer._get_eval_cache`` (model/BaseRecommender.py:109-151):
instead of one LongTensor of train items per eval user plus a LongTensor column mask, it holds one
CSR (mask_rowptr int64, mask_col int32) over the eval users, one CSR of their ground truth and a
per-item flag byte (bit 0 = cold item, bit 1 = warm item).  ``FullRankScorer`` replaces the body of
``_evaluate`` (:170-182): ``batch_predict`` + mask writes + ``torch.topk``.
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import ops

FLAG_COLD = 1   # item listed in data.mapped_cold_item_idx
FLAG_WARM = 2   # item listed in data.mapped_warm_item_idx
GROUP_UNFLAGGED = -1   # item group: rows of the item table listed in neither


def flag_exclude_for(cold_object: str, data_type: str) -> int:
    """Which item flag the reference masks (model/BaseRecommender.py:130-143): item-cold runs mask
    cold items in the 'warm' setting and warm items in the 'cold' setting; nothing otherwise."""
    if cold_object != 'item':
        return 0
    return {'warm': FLAG_COLD, 'cold': FLAG_WARM}.get(data_type, 0)


def item_flags_from(data, device) -> torch.Tensor:
    flags = np.zeros(int(data.item_num), dtype=np.uint8)
    flags[np.asarray(data.mapped_cold_item_idx, dtype=np.int64)] |= FLAG_COLD
    flags[np.asarray(data.mapped_warm_item_idx, dtype=np.int64)] |= FLAG_WARM
    return torch.from_numpy(flags).to(device)


def _csr_from_lists(rows: Sequence[np.ndarray]):
    lens = np.fromiter((len(r) for r in rows), dtype=np.int64, count=len(rows))
    rowptr = np.zeros(len(rows) + 1, dtype=np.int64)
    np.cumsum(lens, out=rowptr[1:])
    col = np.concatenate(rows).astype(np.int32) if len(rows) and rowptr[-1] else np.zeros(0, dtype=np.int32)
    return rowptr, col


def _csr_from_dicts(rows: Sequence[Dict], imap: Dict):
    """CSR of dense item ids, ascending within a row, from one {raw_item: rating} dict per row: one pass over the
    flattened keys through ``imap`` and one lexsort, instead of a numpy sort per user."""
    lens = np.fromiter((len(r) for r in rows), dtype=np.int64, count=len(rows))
    rowptr = np.zeros(len(rows) + 1, dtype=np.int64)
    np.cumsum(lens, out=rowptr[1:])
    total = int(rowptr[-1])
    try:
        flat = np.fromiter(map(imap.__getitem__, itertools.chain.from_iterable(rows)), dtype=np.int64, count=total)
    except KeyError as e:              # util/databuilder.py:283-287
        raise Exception(f"item {e.args[0]} not in current id table")
    rowid = np.repeat(np.arange(len(rows), dtype=np.int64), lens)
    return rowptr, flat[np.lexsort((flat, rowid))].astype(np.int32)


@dataclass
class EvalPlan:
    users: list                    # raw eval-user ids in ground-truth dict order (BaseRecommender.py:115)
    user_ids: torch.Tensor         # int32 [n_q] dense user ids
    mask_rowptr: torch.Tensor      # int64 [n_q+1]   train items of each eval user (:117-128)
    mask_col: torch.Tensor         # int32, ascending within a row
    gt_rowptr: torch.Tensor        # int64 [n_q+1]   ground-truth items, dense ids, ascending within a row
    gt_col: torch.Tensor
    flag_exclude: int              # 0 | FLAG_COLD | FLAG_WARM (:130-143)

    @property
    def n_q(self) -> int:
        return self.user_ids.numel()

    @classmethod
    def from_data(cls, data, data_set: Dict, data_type: str, cold_object: str, device) -> "EvalPlan":
        """From a ColdStartDataBuilder-like object (``user``/``item`` maps, ``training_set_u``)."""
        users = list(data_set.keys())
        umap, imap = data.user, data.item
        try:
            uids = np.fromiter((umap[u] for u in users), dtype=np.int32, count=len(users))
        except KeyError as e:          # util/databuilder.py:289-296
            raise Exception(f"user {e.args[0]} not in current id table")
        tr = data.training_set_u
        empty = {}
        mrp, mc = _csr_from_dicts([tr[u] if u in tr else empty for u in users], imap)      # `in` first: never grows a defaultdict
        grp, gc = _csr_from_dicts([data_set[u] for u in users], imap)
        t = lambda a: torch.from_numpy(a).to(device)
        return cls(users, t(uids), t(mrp), t(mc), t(grp), t(gc), flag_exclude_for(cold_object, data_type))

    def slice(self, lo: int, hi: int) -> "EvalPlan":
        """Eval users [lo, hi) as a plan of their own (user-sharded scoring: every GPU ranks its slice of the users)."""
        def rows(rowptr, col):
            b, e = int(rowptr[lo]), int(rowptr[hi])
            return (rowptr[lo:hi + 1] - b).contiguous(), col[b:e].contiguous()
        mrp, mc = rows(self.mask_rowptr, self.mask_col)
        grp, gc = rows(self.gt_rowptr, self.gt_col)
        return EvalPlan(None if self.users is None else self.users[lo:hi], self.user_ids[lo:hi].contiguous(), mrp, mc, grp, gc,
                        self.flag_exclude)

    @classmethod
    def from_arrays(cls, user_ids, mask_rowptr, mask_col, gt_rowptr, gt_col, flag_exclude=0, users=None) -> "EvalPlan":
        return cls(users if users is not None else None, user_ids, mask_rowptr, mask_col, gt_rowptr, gt_col, int(flag_exclude))


class FullRankScorer:
    """score -> train mask -> flag mask -> top-K for every eval user against the whole catalogue.

    ``tables`` generalises ``batch_predict``: a list of (user_table, item_table, item_group) where
    item_group is None (all items), FLAG_WARM / FLAG_COLD (only items carrying that flag) or
    GROUP_UNFLAGGED (items carrying neither):
      - the canonical ``user_emb[users] @ item_emb.T`` (model/MF.py:58-63)   -> [(U, I, None)]
      - ALDI's dual user tables (model/ALDI.py:149-160)                     -> [(Uw, I, FLAG_WARM), (Uc, I, FLAG_COLD)]
      - VBPR/AMR's sum of two products (model/VBPR.py:68-75)                 -> [([P|P2], [Q|Q2], None)]
    Each table pair is swept by the fused kernel; several lists are merged by (score desc, id asc).
    Items excluded by the warm/cold column mask are compacted away before the sweep when that is
    cheaper; lists that come up short because of it are completed with masked ids at -1e9, as the
    reference's top-K would show.
    """

    def __init__(self, K: int, precision: int = ops.SCORE_TF32_CHECKED):
        self.K, self.precision = int(K), precision
        self._gid_cache = {}                         # (flags identity, group, excl) -> kept item ids; never table contents
        self._remap_cache = {}                       # (plan identity, selection) -> train-mask CSR in rows of the compacted table
        self.n_refined: List[torch.Tensor] = []      # device counters of the last topk() call

    def _kept_ids(self, item_flags: torch.Tensor, keep_mask_fn, key) -> torch.Tensor:
        """Ids of the items kept by a flag selection.  Only this id list is cached (a function of the flag bytes alone);
        table rows are gathered afresh on EVERY call and never cached: trainers rebind ``item_emb`` to a new tensor every
        epoch and the caching allocator hands the same address (and ``_version`` 0) back two epochs later, so any
        (address, version) key would silently return an old epoch's rows."""
        ck = (item_flags.data_ptr(), item_flags._version, item_flags.numel(), key)
        hit = self._gid_cache.get(ck)
        if hit is None or hit[0] is not item_flags:  # the cached entry keeps the flag tensor alive: its address cannot be reused
            gids = torch.nonzero(keep_mask_fn(), as_tuple=False).flatten().to(torch.int32)
            if len(self._gid_cache) >= 16:
                self._gid_cache.clear()
            hit = (item_flags, gids)
            self._gid_cache[ck] = hit
        return hit[1]

    def _mask_in_rows_of(self, plan: EvalPlan, gids: torch.Tensor, key):
        """The plan's train-mask CSR re-expressed in ROW NUMBERS of a compacted item table (``gids`` = the kept item ids,
        ascending): entries whose item was compacted away are dropped, the others become their row number, order kept.
        The sweep then sees a contiguous catalogue 0 .. len(gids)-1 — its fast mask path — instead of locating every train
        item inside every tile of an id list (r02: 0.40-0.45 of the tensor peak on that path against 0.99); the ranked row
        numbers are mapped back through ``gids`` afterwards (row order = id order, so ties break the same way).
        Memoised per (plan, selection): plans are memoised per eval set, the selection is a function of the flag bytes."""
        ck = (id(plan), key)
        hit = self._remap_cache.get(ck)
        # an entry is valid only for the very plan and kept-id tensors it was built from (both kept alive by the entry, so neither
        # identity can be recycled): a rebuilt id list — new flags — can never meet an old remap
        if hit is None or hit[0] is not plan or hit[3] is not gids:
            col = plan.mask_col
            pos = torch.searchsorted(gids, col)
            valid = gids[pos.clamp(max=gids.numel() - 1)] == col
            lens = plan.mask_rowptr[1:] - plan.mask_rowptr[:-1]
            rowid = torch.repeat_interleave(torch.arange(plan.n_q, device=col.device), lens)
            rowptr = torch.zeros(plan.n_q + 1, dtype=torch.int64, device=col.device)
            rowptr[1:] = torch.cumsum(torch.bincount(rowid[valid], minlength=plan.n_q), 0)
            if len(self._remap_cache) >= 16:
                self._remap_cache.clear()
            hit = (plan, rowptr, pos[valid].to(torch.int32).contiguous(), gids)
            self._remap_cache[ck] = hit
        return hit[1], hit[2]

    def topk(self, tables, plan: EvalPlan, item_flags: Optional[torch.Tensor] = None):
        K, excl = self.K, plan.flag_exclude
        if excl and item_flags is None:
            raise ValueError("this plan masks warm/cold items: item_flags is required")
        n_items_total = tables[0][1].shape[0]
        if n_items_total < K:
            raise ValueError(f"top-{K} over {n_items_total} items (torch.topk raises here as well)")
        lists, compacted = [], False
        self.n_refined = []
        for user_tab, item_tab, group in tables:
            tab, gids, kflags, kexcl, sel = item_tab, None, None, 0, None
            if group is not None:
                if item_flags is None:
                    raise ValueError("item groups need item_flags")
                def keep(group=group):
                    k = (item_flags & (FLAG_WARM | FLAG_COLD)) == 0 if group == GROUP_UNFLAGGED else (item_flags & group) != 0
                    return k & ((item_flags & excl) == 0) if excl else k
                sel = (group, excl)
                gids = self._kept_ids(item_flags, keep, sel)
                tab, compacted = ops.gather_rows(item_tab, gids), True
            elif excl:
                sel = (0, excl)
                gids = self._kept_ids(item_flags, lambda: (item_flags & excl) == 0, sel)
                if gids.numel() < 0.9 * item_tab.shape[0]:         # skipping flagged items outright is cheaper
                    tab, compacted = ops.gather_rows(item_tab, gids), True
                else:                                              # few items flagged: mask them inside the kernel instead
                    gids, kflags, kexcl = None, item_flags, excl
            if tab.shape[0] == 0:
                continue
            if gids is not None:       # compacted table: sweep it as a contiguous catalogue of row numbers, map the rows back to ids
                if plan.mask_col.numel():
                    mrp, mcol = self._mask_in_rows_of(plan, gids, sel)
                else:
                    mrp, mcol = plan.mask_rowptr, plan.mask_col
                s, i, nref = ops.score_topk(user_tab, tab, K, user_ids=plan.user_ids, mask_rowptr=mrp, mask_col=mcol, precision=self.precision)
                i = torch.where(i >= 0, gids[i.clamp(min=0).long()], i)
            else:
                s, i, nref = ops.score_topk(user_tab, tab, K, user_ids=plan.user_ids, mask_rowptr=plan.mask_rowptr,
                                            mask_col=plan.mask_col, item_flags=kflags, flag_exclude=kexcl, precision=self.precision)
            self.n_refined.append(nref)
            lists.append((s, i))
        if not lists:
            s = torch.full((plan.n_q, K), float("-inf"), dtype=torch.float32, device=plan.user_ids.device)
            i = torch.full((plan.n_q, K), -1, dtype=torch.int32, device=plan.user_ids.device)
        elif len(lists) == 1:
            s, i = lists[0]
        else:
            s, i = ops.topk_merge(torch.stack([l[0] for l in lists]), torch.stack([l[1] for l in lists]))
        if compacted:
            ops.fill_masked(s, i, n_items_total, item_flags=item_flags if excl else None, flag_exclude=excl,
                            mask_rowptr=plan.mask_rowptr, mask_col=plan.mask_col)
        return s, i


class HostBatchEvaluator:
    """End-to-end evaluation of one eval batch whose plan lives in HOST memory (the shape of the
    reference's ``_evaluate`` call: users + their train items + ground truth come from Python-side
    structures, model tables stay on the device): pinned host -> device copies of the plan, fused
    scoring, device metrics, then device -> host copies of the top-K lists and the metric sums.
    ``scorer`` is a ``coldrec_b200.dist.GridShardedFullRankScorer`` (world size 1 included).

    Each rank stages only what it sweeps: ``pin`` keeps the rows of its USER GROUP (all users when the catalogue is split
    over every rank, 1/groups of them in a grid layout) — at 8 GPUs in the 4 x 2 grid half of the 279 MB the first version
    copied to every rank.  Copies go through a copy stream into one of two device slots, so the plan of step k+1
    (``next_host_plan``) crosses PCIe while step k is swept."""

    def __init__(self, scorer, Ns: Sequence[int], n_q: int, max_mask_nnz: int, max_gt_nnz: int, device):
        self.scorer, self.Ns, self.n_q, self.device = scorer, list(Ns), n_q, device
        self.g_lo, self.g_hi = scorer.group_slice(n_q) if hasattr(scorer, "group_slice") else (0, n_q)
        n_g = self.g_hi - self.g_lo
        e = lambda n, dt: torch.empty(n, dtype=dt, device=device)
        self.slots = [dict(user_ids=e(n_g, torch.int32), mask_rowptr=e(n_g + 1, torch.int64), mask_col=e(max_mask_nnz, torch.int32),
                           gt_rowptr=e(n_g + 1, torch.int64), gt_col=e(max_gt_nnz, torch.int32)) for _ in range(2)]
        self.copy_stream = torch.cuda.Stream(device=device)
        self.copied = [torch.cuda.Event() for _ in range(2)]        # slot filled
        self.released = [torch.cuda.Event() for _ in range(2)]      # slot no longer read by the sweep
        self._staged = [None, None]
        self._turn = 0
        lo, hi = scorer.user_slice(n_q) if scorer.world > 1 else (0, n_q)
        self.out_ids = torch.empty((hi - lo, scorer.K), dtype=torch.int32).pin_memory()
        self.out_scores = torch.empty((hi - lo, scorer.K), dtype=torch.float32).pin_memory()
        self.h2d_bytes = 0
        self.d2h_bytes = (hi - lo) * scorer.K * 8 + len(self.Ns) * 6 * 8

    def pin(self, plan: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        """This rank's user group of a host plan, in pinned memory (row pointers rebased to the group)."""
        lo, hi = self.g_lo, self.g_hi
        out = {"user_ids": plan["user_ids"][lo:hi]}
        for rp, col in (("mask_rowptr", "mask_col"), ("gt_rowptr", "gt_col")):
            b, e = int(plan[rp][lo]), int(plan[rp][hi])
            out[rp], out[col] = plan[rp][lo:hi + 1] - b, plan[col][b:e]
        return {k: v.detach().cpu().contiguous().pin_memory() for k, v in out.items()}

    def _stage(self, slot: int, host_plan: Dict[str, torch.Tensor]) -> None:
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.released[slot])
            for k, h in host_plan.items():
                self.slots[slot][k][:h.numel()].copy_(h, non_blocking=True)
            self.copied[slot].record(self.copy_stream)
        self._staged[slot] = host_plan

    def run(self, user_tab, item_shard, item_begin: int, host_plan: Dict[str, torch.Tensor], next_host_plan=None):
        slot = self._turn
        if self._staged[slot] is not host_plan:
            self._stage(slot, host_plan)
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(self.copied[slot])
        if next_host_plan is not None:                       # the next step's plan crosses PCIe while this one is swept
            self._stage(slot ^ 1, next_host_plan)
        views = {k: self.slots[slot][k][:h.numel()] for k, h in host_plan.items()}
        self.h2d_bytes = sum(h.numel() * h.element_size() for h in host_plan.values())
        plan = EvalPlan.from_arrays(**views)
        grouped = hasattr(self.scorer, "group_slice")
        s, i = self.scorer.topk(user_tab, item_shard, item_begin, plan, **({"n_q_total": self.n_q} if grouped else {}))
        self.out_ids.copy_(i, non_blocking=True)
        self.out_scores.copy_(s, non_blocking=True)
        perf = self.scorer.metrics(i, plan, self.Ns, rounded=True, **({"n_q_total": self.n_q} if grouped else {}))   # .cpu() of the sums: the per-step sync
        self.released[slot].record(cur)
        self._staged[slot] = None
        self._turn ^= 1
        return perf
