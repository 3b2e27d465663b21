"""Training-side callers of the propagation path: sampler, fused BPR step, Adam.

``PairwiseSampler`` replaces ``next_batch_pairwise`` (util/utils.py:123-157) and ``BprTrainStep`` the loop body of
``LightGCN.train`` (model/LightGCN.py:21-28; with ``n_layers=0`` it is ``MF.train``, model/MF.py:19-27):

    for user_idx, pos_idx, neg_idx in sampler.epoch(e, batch_size):      # next_batch_pairwise(self.data, self.batch_size)
        loss = step(user_idx, pos_idx, neg_idx)                          # model() -> gathers -> bpr+l2 -> backward -> optimizer.step()
    user_emb, item_emb = step.embeddings()                               # model() under torch.no_grad()

One step = forward propagation (L SpMM launches, layer mean fused), ``cr_bpr_fwd_bwd_f32`` (loss + the scatter-add
gradients of the three row gathers), backward propagation — the SAME SpMM kernel applied to the gradient table, because
the normalised adjacency is symmetric (util/databuilder.py:236-248) and d/dE0 mean_k(A^k E0) = mean_k(A^k) —
and ``cr_adam_step_f32`` on the concatenated parameter table.  No autograd graph, no (N, L+1, d) stack, no COO tensor.
"""
from __future__ import annotations

from typing import Iterator, Optional, Tuple

import numpy as np
import torch

from . import ops
from .graph import CsrGraph, PropagationBuffers, propagate_table


class PairwiseSampler:
    """Device-resident training pairs + train CSR; ``epoch(e, batch_size)`` yields (user, pos, neg) int32 tensors.

    The item table negatives are drawn from has ``n_items`` rows = ``len(data.item)`` in the reference
    (``item_list = list(data.item.keys())``, util/utils.py:129): every mapped item, not only training items.
    """

    def __init__(self, pair_user: torch.Tensor, pair_item: torch.Tensor, n_users: int, n_items: int, seed: int = 2024):
        if not (pair_user.is_cuda and pair_item.is_cuda):
            raise ValueError("PairwiseSampler samples on the device: pass CUDA tensors (coldrec_b200 has no CPU path)")
        self.pair_user = pair_user.to(torch.int32).contiguous()
        self.pair_item = pair_item.to(torch.int32).contiguous()
        self.n_users, self.n_items, self.seed = int(n_users), int(n_items), int(seed)
        dev = pair_user.device
        # the train CSR (sorted, duplicate pairs collapsed): the same structure the scorer masks with
        keys = torch.unique(self.pair_user.to(torch.int64) * self.n_items + self.pair_item.to(torch.int64), sorted=True)
        rows = torch.div(keys, self.n_items, rounding_mode="floor")
        self.train_col = (keys - rows * self.n_items).to(torch.int32).contiguous()
        self.train_rowptr = torch.zeros(self.n_users + 1, dtype=torch.int64, device=dev)
        self.train_rowptr[1:] = torch.cumsum(torch.bincount(rows, minlength=self.n_users), 0)
        self.n_exhausted = torch.zeros(1, dtype=torch.int32, device=dev)

    @classmethod
    def from_data(cls, data, device, seed: int = 2024) -> "PairwiseSampler":
        """From a ColdStartDataBuilder-like object (``training_data`` triples + ``user`` / ``item`` id maps)."""
        u = np.fromiter((data.user[p[0]] for p in data.training_data), dtype=np.int32, count=len(data.training_data))
        i = np.fromiter((data.item[p[1]] for p in data.training_data), dtype=np.int32, count=len(data.training_data))
        return cls(torch.from_numpy(u).to(device), torch.from_numpy(i).to(device), max(int(data.user_num), len(data.user)),
                   len(data.item), seed)

    @property
    def n_pairs(self) -> int:
        return self.pair_user.numel()

    def batch(self, epoch: int, begin: int, count: int, out: Optional[torch.Tensor] = None):
        return ops.sample_pairwise(self.pair_user, self.pair_item, self.train_rowptr, self.train_col, self.n_items, self.seed, epoch,
                                   begin, count, out=out, n_exhausted=self.n_exhausted)

    def epoch(self, epoch: int, batch_size: int) -> Iterator[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]]:
        """Batches of one pass over the shuffled training pairs (the last one may be short, util/utils.py:131-135)."""
        ptr = 0
        while ptr < self.n_pairs:
            count = min(batch_size, self.n_pairs - ptr)
            yield self.batch(epoch, ptr, count)
            ptr += count


class BprTrainStep:
    """Parameters, Adam state and scratch of BPR training over a (propagated) embedding table.

    ``graph`` is the normalised bipartite adjacency (``CsrGraph``), or None / ``n_layers=0`` for plain MF.
    Parameters live in ONE (n_users + n_items, d) table ``self.ego`` (``user_emb`` / ``item_emb`` are views), so the
    propagation needs no ``torch.cat`` per step and Adam is one launch.
    """

    def __init__(self, graph: Optional[CsrGraph], user_emb: torch.Tensor, item_emb: torch.Tensor, n_layers: int, lr: float,
                 reg: float, betas=(0.9, 0.999), eps: float = 1e-8, include_ego: bool = True):
        if not (user_emb.is_cuda and item_emb.is_cuda):
            raise ValueError("BprTrainStep trains on the device: pass CUDA tensors (coldrec_b200 has no CPU path)")
        if n_layers > 0 and graph is None:
            raise ValueError("n_layers > 0 needs the adjacency")
        self.graph, self.n_layers, self.lr, self.reg, self.betas, self.eps = graph, int(n_layers), lr, reg, betas, eps
        self.include_ego = include_ego
        self.n_users, self.n_items, self.d = user_emb.shape[0], item_emb.shape[0], user_emb.shape[1]
        dev = user_emb.device
        self.ego = torch.cat([user_emb.detach().to(torch.float32), item_emb.detach().to(torch.float32)], 0).contiguous()
        n = self.ego.shape[0]
        self.exp_avg = torch.zeros_like(self.ego)
        self.exp_avg_sq = torch.zeros_like(self.ego)
        self.grad_out = torch.zeros_like(self.ego)          # d loss / d (propagated table)
        self.fwd = PropagationBuffers(n, self.d, dev) if self.n_layers > 0 else None
        self.bwd = PropagationBuffers(n, self.d, dev) if self.n_layers > 0 else None
        self.loss = torch.zeros(4, dtype=torch.float32, device=dev)
        self._ws = None
        self.steps = 0

    # views ------------------------------------------------------------------------------------------
    @property
    def user_emb(self) -> torch.Tensor:
        return self.ego[:self.n_users]

    @property
    def item_emb(self) -> torch.Tensor:
        return self.ego[self.n_users:]

    def _forward(self) -> torch.Tensor:
        if self.n_layers == 0:
            return self.ego
        return propagate_table(self.graph, self.ego, self.n_layers, self.include_ego, buffers=self.fwd)

    def embeddings(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """``model()`` in eval mode (model/LightGCN.py:34): the propagated user / item tables (fresh tensors)."""
        out = self._forward().clone()
        return out[:self.n_users], out[self.n_users:]

    def gradients(self, u_idx, i_idx, j_idx) -> torch.Tensor:
        """loss.backward() of one batch: returns d loss / d ego (N, d); ``self.loss`` holds (total, bpr, reg)."""
        table = self._forward()
        self.grad_out.zero_()
        B = u_idx.numel()
        need = ops._lib.load().cr_bpr_workspace_bytes(B)
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.ego.device)
        ops.bpr_fwd_bwd(table[:self.n_users], table[self.n_users:], u_idx, i_idx, j_idx, self.reg, self.grad_out[:self.n_users],
                        self.grad_out[self.n_users:], loss=self.loss, workspace=self._ws)
        if self.n_layers == 0:
            return self.grad_out
        # backward of mean_k(A^k . ego): A is symmetric, so the same propagation applied to the output gradient
        return propagate_table(self.graph, self.grad_out, self.n_layers, self.include_ego, buffers=self.bwd)

    def step(self, u_idx: torch.Tensor, i_idx: torch.Tensor, j_idx: torch.Tensor) -> torch.Tensor:
        """One optimisation step; returns the device loss tensor [total, bpr, reg, 0] (no host sync)."""
        grad = self.gradients(u_idx, i_idx, j_idx)
        self.steps += 1
        ops.adam_step(self.ego, grad, self.exp_avg, self.exp_avg_sq, self.steps, self.lr, self.betas, self.eps)
        return self.loss

    # ---- CUDA-graph replay for launch-bound graphs ----------------------------------------------------
    def capture(self, batch_size: int) -> "CapturedTrainStep":
        """Capture forward + BPR + backward + Adam for batches of exactly ``batch_size`` samples into one CUDA graph.
        On dataset-sized graphs (CiteULike: 22k nodes, 0.3M nonzeros) a step is ~25 kernels of a few microseconds each
        and launch latency dominates; the graph replays them with one launch.  The step-dependent Adam factors are
        read from device memory, refreshed (8 bytes) before each replay."""
        return CapturedTrainStep(self, batch_size)


class CapturedTrainStep:
    """``BprTrainStep.step`` for a fixed batch size as a CUDA graph.  ``idx`` is the static int32 [3, B] batch buffer —
    let the sampler write into it (``sampler.batch(epoch, begin, B, out=cap.idx)``) or pass tensors to ``__call__``."""

    _RING = 256      # pinned slots for the per-step Adam factors: a slot is rewritten only after a stream sync

    def __init__(self, step: BprTrainStep, batch_size: int):
        self.step, self.B = step, int(batch_size)
        self._calls = 0
        dev = step.ego.device
        self.idx = torch.zeros((3, self.B), dtype=torch.int32, device=dev)
        self.scalars = torch.zeros(2, dtype=torch.float32, device=dev)
        self._host = torch.zeros((self._RING, 2), dtype=torch.float32).pin_memory()
        # warm up on a side stream (allocations, SpMM plans, function attributes), restoring the state it touches
        saved = [t.clone() for t in (step.ego, step.exp_avg, step.exp_avg_sq)]
        s = torch.cuda.Stream(device=dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s):
            self._body()
        torch.cuda.current_stream(dev).wait_stream(s)
        torch.cuda.synchronize(dev)
        for t, v in zip((step.ego, step.exp_avg, step.exp_avg_sq), saved):
            t.copy_(v)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._body()

    def _body(self):
        st = self.step
        grad = st.gradients(self.idx[0], self.idx[1], self.idx[2])
        ops.adam_step(st.ego, grad, st.exp_avg, st.exp_avg_sq, 1, st.lr, st.betas, st.eps, dev_scalars=self.scalars)

    def __call__(self, u_idx=None, i_idx=None, j_idx=None) -> torch.Tensor:
        st = self.step
        if u_idx is not None:
            if u_idx.numel() != self.B:
                raise ValueError(f"captured for batches of {self.B}, got {u_idx.numel()} (run the short last batch through step())")
            self.idx[0].copy_(u_idx); self.idx[1].copy_(i_idx); self.idx[2].copy_(j_idx)
        st.steps += 1
        slot = self._calls % self._RING
        if slot == 0 and self._calls:
            torch.cuda.current_stream(st.ego.device).synchronize()      # the copies that read the ring have all run
        self._calls += 1
        a, b = ops.adam_scalars(st.steps, st.lr, st.betas)
        self._host[slot, 0], self._host[slot, 1] = a, b
        self.scalars.copy_(self._host[slot], non_blocking=True)
        self.graph.replay()
        return st.loss
