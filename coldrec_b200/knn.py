"""Content kNN through the fused scorer.

``faiss.IndexFlatIP(d).add(value); index.search(query, k)`` (model/KNN.py:63-77) is a brute-force inner-product top-k —
exactly what the fused scorer computes with the content vectors as the two tables, so the KNN model needs neither faiss
nor a (n_query, n_value) similarity matrix.  FSGNN's chunked cosine kNN (model/FSGNN.py:106-152) is the same sweep over
row-normalised features with each row's own id as its one-entry mask row.  The cold-row generator of the KNN model,
``mean(emb_table[neighbours], dim=1)`` (model/KNN.py:79-88), is the SpMM kernel with all-ones values and the division
in its epilogue.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import numpy as np
import torch

from . import ops


def _table(x, device) -> torch.Tensor:
    """fp32 device table with the width padded to a multiple of 4 (zero columns do not change inner products)."""
    t = torch.as_tensor(np.ascontiguousarray(x) if isinstance(x, np.ndarray) else x)
    t = t.to(device=device, dtype=torch.float32)
    pad = (-t.shape[1]) % 4
    if pad:
        t = torch.nn.functional.pad(t, (0, pad))
    return t.contiguous()


def knn_inner_product(query, value, k: int, device="cuda:0", exclude_self: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """Top-k rows of ``value`` by inner product for every row of ``query``: (scores [n_q,k] fp32, ids [n_q,k] int32),
    sorted descending like ``IndexFlatIP.search``.  ``exclude_self`` masks value row j for query row j (kNN graphs)."""
    q, v = _table(query, device), _table(value, device)
    if q.shape[1] != v.shape[1]:
        raise ValueError(f"query width {q.shape[1]} != value width {v.shape[1]}")
    if k > v.shape[0] - (1 if exclude_self else 0):
        raise ValueError(f"k={k} exceeds the number of candidates")
    if q.shape[1] > 128 and not os.environ.get("CR_KNN_SIMT"):
        return _knn_wide(q, v, k, exclude_self)
    rowptr = col = None
    if exclude_self:
        rowptr = torch.arange(q.shape[0] + 1, dtype=torch.int64, device=q.device)
        col = torch.arange(q.shape[0], dtype=torch.int32, device=q.device)
    # widths <= 128: the tcgen05 sweep (TF32-checked: exact fp32 scores of a proven top-k; narrower tables zero-padded to 64 / 128)
    s, i, _ = ops.score_topk(q, v, k, mask_rowptr=rowptr, mask_col=col, precision=ops.SCORE_TF32_CHECKED)
    return s, i


def _knn_wide(q: torch.Tensor, v: torch.Tensor, k: int, exclude_self: bool, block_bytes: int = 512 << 20):
    """Content tables wider than the fused sweep's TMEM-resident query tiles (CiteULike 300, XING 2,738): inner products of a
    block of queries against every value on the tensor cores at fp32 accuracy (``cr_linear_act_tc_f32``, the value table as
    the weight matrix — 3xTF32 with K-block drains), then the row-wise top-k kernel; the (queries x values) block is the only
    score storage and is reused from block to block."""
    vs = ops.split_tf32(v)
    n_q, n_v = q.shape[0], v.shape[0]
    rows = max(128, min(n_q, block_bytes // (4 * max(n_v, 1)) // 128 * 128))
    S = torch.empty((min(rows, n_q), n_v), dtype=torch.float32, device=q.device)
    out_s = torch.empty((n_q, k), dtype=torch.float32, device=q.device)
    out_i = torch.empty((n_q, k), dtype=torch.int32, device=q.device)
    for lo in range(0, n_q, rows):
        hi = min(n_q, lo + rows)
        ops.linear_act_tc(ops.split_tf32(q[lo:hi]), vs, None, out=S[:hi - lo])     # (pre-split rows: 3.0 vs 4.2 ms with the in-kernel split, XING shape)
        ex = torch.arange(lo, hi, dtype=torch.int32, device=q.device) if exclude_self else None
        out_s[lo:hi], out_i[lo:hi] = ops.topk_rows(S[:hi - lo], k, exclude_col=ex)
    return out_s, out_i


def precompute_knn_neighbors(data, cold_object: str, knn_num: int, device="cuda:0") -> np.ndarray:
    """``KNN._precompute_knn_neighbors`` (model/KNN.py:63-77): for every cold item (user) the mapped ids of its
    ``knn_num`` nearest warm items (users) by content inner product."""
    if cold_object == 'item':
        content, cold, warm = data.mapped_item_content, data.mapped_cold_item_idx, data.mapped_warm_item_idx
    else:
        content, cold, warm = data.mapped_user_content, data.mapped_cold_user_idx, data.mapped_warm_user_idx
    cold, warm = np.asarray(cold, dtype=np.int64), np.asarray(warm, dtype=np.int64)
    _, ids = knn_inner_product(content[cold], content[warm], knn_num, device)
    return warm[ids.cpu().numpy().astype(np.int64)]


def knn_generate(emb_table: torch.Tensor, neighbor_ids: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``KNN.knn_search`` (model/KNN.py:79-88): ``mean(emb_table[neighbor_ids], dim=1)`` for an (n, k) id matrix."""
    n, k = neighbor_ids.shape
    dev = emb_table.device
    rowptr = torch.arange(0, (n + 1) * k, k, dtype=torch.int64, device=dev)
    col = neighbor_ids.to(device=dev, dtype=torch.int32).contiguous().view(-1)
    if out is None:
        out = torch.empty((n, emb_table.shape[1]), dtype=torch.float32, device=dev)
    ops.spmm(rowptr, col, None, emb_table.contiguous(), acc=out, acc_beta=0.0, acc_div=float(k))
    return out


def cosine_knn_graph(feat, k: int, device="cuda:0"):
    """Neighbour lists of FSGNN's ``_cosine_knn_graph`` (model/FSGNN.py:106-152) before its scipy symmetrisation:
    (similarity [n,k], neighbour ids [n,k]) of the k most cosine-similar OTHER rows."""
    x = torch.as_tensor(np.asarray(feat)).to(device=device, dtype=torch.float64)
    x = (x / torch.clamp(torch.linalg.norm(x, dim=1, keepdim=True), min=1e-12)).to(torch.float32)       # :118-121
    k_eff = min(int(k), x.shape[0] - 1)
    return knn_inner_product(x, x, k_eff, device, exclude_self=True)
