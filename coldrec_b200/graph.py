"""Graph side of the hot path: device-resident CSR adjacency and LightGCN-family propagation.

Mirrors ``TorchGraphInterface.convert_sparse_mat_to_tensor`` (util/databuilder.py:953-962) and the
propagation loops of ``LGCN_Encoder.forward`` (model/LightGCN.py:86-96), ``SimGCL_Encoder.forward``
(model/SimGCL.py:101-113) and ``NGCF_Encoder.forward`` (model/NGCF.py:90-104): where the reference
keeps an int64 COO tensor and calls ``torch.sparse.mm`` + ``stack``/``mean``, this keeps an int32 CSR
and calls the fused SpMM kernel (``cr_spmm_csr_f32``) with the layer mean folded into its epilogue.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np
import torch

from . import ops


class CsrGraph:
    """A (weighted) adjacency in CSR on one device: rowptr int64 [n_rows+1], col int32, val fp32.

    ``n_cols`` is the number of rows of the gather source; for a row-partitioned graph ``n_rows`` is
    the local block and ``row_begin`` its first global row.

    Single-stream object: the long-row plan cached per width (``plan(d)``) also holds the partial-sum scratch of the split
    path, so two SpMMs on the SAME graph must not run concurrently on different streams (e.g. an eval-time propagate
    overlapping a captured training step) — use one ``CsrGraph`` (one plan) per stream; the CSR arrays themselves can be shared.
    """

    def __init__(self, rowptr: torch.Tensor, col: torch.Tensor, val: Optional[torch.Tensor], n_cols: int,
                 row_begin: int = 0):
        self.rowptr, self.col, self.val = rowptr, col, val
        self.n_rows, self.n_cols, self.row_begin = rowptr.numel() - 1, int(n_cols), int(row_begin)
        self.nnz = col.numel()
        self._plans = {}

    @classmethod
    def from_scipy(cls, mat, device) -> "CsrGraph":
        """The drop-in for ``TorchGraphInterface.convert_sparse_mat_to_tensor(data.norm_adj).to(device)``."""
        csr = mat.tocsr()
        csr.sum_duplicates()
        csr.sort_indices()
        if csr.shape[1] > np.iinfo(np.int32).max:
            raise ValueError("column ids do not fit int32")
        return cls(torch.from_numpy(csr.indptr.astype(np.int64)).to(device),
                   torch.from_numpy(csr.indices.astype(np.int32)).to(device),
                   torch.from_numpy(csr.data.astype(np.float32)).to(device), csr.shape[1])

    @classmethod
    def from_torch_sparse(cls, coo: torch.Tensor, device=None) -> "CsrGraph":
        """From the reference's coalesced COO tensor (what its encoders hold as ``sparse_norm_adj``)."""
        coo = coo.coalesce()
        device = device or coo.device
        idx, val = coo.indices(), coo.values()
        counts = torch.bincount(idx[0], minlength=coo.shape[0])
        rowptr = torch.zeros(coo.shape[0] + 1, dtype=torch.int64, device=idx.device)
        rowptr[1:] = torch.cumsum(counts, 0)
        return cls(rowptr.to(device), idx[1].to(torch.int32).to(device), val.to(torch.float32).to(device), coo.shape[1])

    def row_block(self, begin: int, end: int) -> "CsrGraph":
        """Rows [begin, end) as a local block (row-partitioned multi-GPU SpMM)."""
        lo, hi = int(self.rowptr[begin]), int(self.rowptr[end])
        rp = (self.rowptr[begin:end + 1] - lo).contiguous()
        return CsrGraph(rp, self.col[lo:hi].contiguous(), None if self.val is None else self.val[lo:hi].contiguous(),
                        self.n_cols, row_begin=self.row_begin + begin)

    def plan(self, d: int) -> Optional[torch.Tensor]:
        if d not in self._plans:
            self._plans[d] = ops.spmm_plan(self.rowptr, self.nnz, d)
        return self._plans[d]

    def spmm(self, X, Y=None, acc=None, acc_in=None, acc_beta=1.0, acc_div=1.0):
        """``torch.sparse.mm(adj, X)`` (model/LightGCN.py:90): a new (n_rows, d) tensor unless Y / acc are given."""
        if Y is None and acc is None:
            Y = torch.empty((self.n_rows, X.shape[1]), dtype=torch.float32, device=X.device)
        return ops.spmm(self.rowptr, self.col, self.val, X, Y=Y, acc=acc, acc_in=acc_in, acc_beta=acc_beta, acc_div=acc_div,
                        plan=self.plan(X.shape[1]))


def bipartite_norm_csr(user_idx: torch.Tensor, item_idx: torch.Tensor, user_num: int, item_num: int) -> CsrGraph:
    """Device-side builder of the normalised bipartite adjacency, array in / CSR out.

    Same result as ``create_sparse_complete_bipartite_adjacency`` + ``normalize_graph_mat``
    (util/databuilder.py:220-254): A = [[0,R],[R^T,0]] with duplicate pairs summed, then
    D^-1/2 A D^-1/2 evaluated as (d[r]*a)*d[c] in fp32, zero-degree rows left at 0.  The reference
    builds it from Python lists over every interaction (:225-226); this is the vectorised form for
    graphs of 10^8 edges.  user_idx/item_idx: int64 CUDA tensors of dense ids (one entry per train pair).
    """
    if not (user_idx.is_cuda and item_idx.is_cuda):
        raise ValueError("bipartite_norm_csr builds on the device: pass CUDA index tensors")
    n = user_num + item_num
    u = user_idx.to(torch.int64)
    i = item_idx.to(torch.int64) + user_num
    keys = torch.cat([u * n + i, i * n + u])
    keys, counts = torch.unique(keys, sorted=True, return_counts=True)
    rows, cols = torch.div(keys, n, rounding_mode="floor"), keys % n
    a = counts.to(torch.float32)
    deg = torch.zeros(n, dtype=torch.float32, device=u.device).index_add_(0, rows, a)
    dinv = torch.where(deg != 0, deg.pow(-0.5), torch.zeros_like(deg))
    val = (dinv[rows] * a) * dinv[cols]
    rowptr = torch.zeros(n + 1, dtype=torch.int64, device=u.device)
    rowptr[1:] = torch.cumsum(torch.bincount(rows, minlength=n), 0)
    return CsrGraph(rowptr, cols.to(torch.int32), val, n)


class PropagationBuffers:
    """Scratch of one propagation over an (N, d) table: the layer-mean accumulator and two ping-pong layer tables.
    Reusing one instance across calls (training steps, epochs) keeps the hot loop free of allocations."""

    def __init__(self, n: int, d: int, device, with_ego: bool = False):
        self.acc = torch.empty((n, d), dtype=torch.float32, device=device)
        self.ping = [torch.empty((n, d), dtype=torch.float32, device=device) for _ in range(2)]
        # room for the [U; I] concatenation of ``propagate`` (the torch.cat of LGCN_Encoder.forward, model/LightGCN.py:87)
        self.ego = torch.empty((n, d), dtype=torch.float32, device=device) if with_ego else None


def propagate_table(graph: CsrGraph, ego: torch.Tensor, n_layers: int, include_ego: bool = True, return_layers: bool = False,
                    buffers: Optional[PropagationBuffers] = None):
    """``propagate`` on the concatenated (N, d) table ``ego = [U; I]``; returns the (N, d) layer mean (a view of
    ``buffers.acc`` when buffers are given) and, with ``return_layers``, the list of layer tables."""
    if n_layers < 1:
        raise ValueError("n_layers must be >= 1")
    if ego.shape[0] != graph.n_cols or graph.n_rows != graph.n_cols:
        raise ValueError(f"adjacency is {graph.n_rows}x{graph.n_cols}, embeddings have {ego.shape[0]} rows")
    count = n_layers + (1 if include_ego else 0)
    acc = buffers.acc if buffers is not None else torch.empty_like(ego)
    layers: List[torch.Tensor] = [ego] if include_ego else []
    x = ego
    bufs = None
    if n_layers > 1 and not return_layers:
        bufs = buffers.ping if buffers is not None else [torch.empty_like(ego), torch.empty_like(ego)]
    for k in range(1, n_layers + 1):
        last = k == n_layers
        need_y = (not last) or return_layers
        y = None
        if need_y:
            y = torch.empty_like(ego) if return_layers else bufs[k & 1]
        first = k == 1
        graph.spmm(x, Y=y, acc=acc, acc_in=(ego if (first and include_ego) else None),
                   acc_beta=(0.0 if (first and not include_ego) else 1.0), acc_div=(float(count) if last else 1.0))
        if return_layers:
            layers.append(y)
        x = y
    return (acc, layers) if return_layers else acc


def propagate(graph: CsrGraph, user_emb: torch.Tensor, item_emb: torch.Tensor, n_layers: int, include_ego: bool = True,
              return_layers: bool = False, buffers: Optional[PropagationBuffers] = None):
    """E0 = [U; I];  E_{k+1} = A.E_k;  result = mean over layers 0..L (LightGCN, model/LightGCN.py:86-96)
    or 1..L (``include_ego=False``: SimGCL/XSimGCL eval path, model/SimGCL.py:101-113).

    One SpMM launch group per layer; the running layer sum and the final division live in the SpMM
    epilogue, so no (N, L+1, d) stack is ever materialised.  ``return_layers`` additionally returns
    [E_0 (if include_ego), E_1, ..., E_L] like NCL's encoder (model/NCL.py:186-196).
    """
    n_u = user_emb.shape[0]
    if buffers is not None and buffers.ego is not None:      # allocation-free call (a trainer evaluates every epoch: 4 x (N, d) otherwise);
        ego = torch.cat([user_emb, item_emb], 0, out=buffers.ego)       # the result is then a view of buffers.acc, valid until the next call
    else:
        ego = torch.cat([user_emb, item_emb], 0).contiguous()
    if return_layers:
        acc, layers = propagate_table(graph, ego, n_layers, include_ego, True, buffers=buffers)
        return acc[:n_u], acc[n_u:], layers
    acc = propagate_table(graph, ego, n_layers, include_ego, buffers=buffers)
    return acc[:n_u], acc[n_u:]


def propagate_frozen_cold(graph: CsrGraph, user_emb: torch.Tensor, item_x: torch.Tensor, n_layers: int,
                          cold_item_idx: torch.Tensor) -> List[torch.Tensor]:
    """CGRC ``_propagate_gprime_frozen_cold`` (model/CGRC.py:76-93): LightGCN convolutions on G' in which the cold item
    rows are put back to their content vector ``item_x[cold]`` after every layer.  Returns the layer list
    [h^(0), ..., h^(L)], each (n_users + n_items, d).  ``cold_item_idx``: dense item ids (any integer dtype, may be empty)."""
    if n_layers < 1:
        raise ValueError("n_layers must be >= 1")
    n_u = user_emb.shape[0]
    item_x = item_x.contiguous()
    cold = cold_item_idx.to(device=item_x.device, dtype=torch.int32).contiguous()
    cold_rows = (cold + n_u).contiguous()
    h = torch.cat([user_emb, item_x], 0).contiguous()
    if h.shape[0] != graph.n_cols or graph.n_rows != graph.n_cols:
        raise ValueError(f"adjacency is {graph.n_rows}x{graph.n_cols}, embeddings have {h.shape[0]} rows")
    out = [h]
    for _ in range(n_layers):
        h = graph.spmm(h)
        if cold.numel():
            ops.copy_rows(item_x, h, src_ids=cold, dst_ids=cold_rows)
        out.append(h)
    return out


def propagate_ngcf(graph: CsrGraph, user_emb, item_emb, W_gc, W_bi):
    """``NGCF_Encoder.forward`` (model/NGCF.py:90-104).  Only the ``torch.sparse.mm`` of :95 is on
    the hot path; the dense d x d transforms and the leaky-relu stay library calls."""
    import torch.nn.functional as F
    n_u = user_emb.shape[0]
    ego = torch.cat([user_emb, item_emb], 0).contiguous()
    layers = [ego]
    for (wg, bg), (wb, bb) in zip(W_gc, W_bi):
        side = graph.spmm(ego, Y=torch.empty_like(ego))
        ego = F.leaky_relu(F.linear(side, wg, bg) + F.linear(ego * side, wb, bb)).contiguous()
        layers.append(ego)
    mean = torch.mean(torch.stack(layers, dim=1), dim=1)
    return mean[:n_u], mean[n_u:]
