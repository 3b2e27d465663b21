"""Tensor-level wrappers over the C ABI.  PyTorch is used only for device memory and streams.

Every function validates device / dtype / contiguity and raises instead of casting silently, then
passes raw device pointers and the current CUDA stream to ``libcoldrec_b200.so``.
"""
from __future__ import annotations

import ctypes
import math
from typing import Optional, Sequence, Tuple

import torch

from . import _lib

__all__ = ["spmm_plan", "spmm", "spmm_bcast", "score_topk", "topk_merge", "fill_masked", "gather_rows", "rank_metrics", "linear_act", "bn_fold",
           "SplitTable", "split_tf32", "tma_rows", "linear_act_tc", "topk_rows",
           "heater_blend", "bpr_fwd_bwd", "adam_step", "sample_pairwise", "SCORE_EXACT_F32", "SCORE_TF32_CHECKED"]

SCORE_EXACT_F32 = _lib.SCORE_EXACT_F32
SCORE_TF32_CHECKED = _lib.SCORE_TF32_CHECKED
_ACTS = {None: _lib.ACT_NONE, "none": _lib.ACT_NONE, "tanh": _lib.ACT_TANH, "leaky_relu": _lib.ACT_LEAKY_RELU}


def _req(t: Optional[torch.Tensor], dtype, name: str, optional=False, contiguous=True):
    if t is None:
        if optional:
            return None
        raise ValueError(f"{name} is required")
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (coldrec_b200 has no CPU path)")
    if t.dtype != dtype:
        raise ValueError(f"{name} must be {dtype}, got {t.dtype} (no silent casts)")
    if contiguous and not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    return t


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream(dev) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _same_device(*ts):
    devs = {t.device for t in ts if t is not None}
    if len(devs) > 1:
        raise ValueError(f"tensors live on different devices: {devs}")
    return next(iter(devs))


# ------------------------------------------------------------------------------------------------ K3
def spmm_plan(rowptr: torch.Tensor, nnz: int, d: int) -> torch.Tensor:
    """Build the long-row split plan for one (rowptr, nnz, d); returns the opaque device buffer."""
    lib = _lib.load()
    rowptr = _req(rowptr, torch.int64, "rowptr")
    n_rows = rowptr.numel() - 1
    nbytes = lib.cr_spmm_plan_bytes(n_rows, nnz, d)
    plan = torch.empty(nbytes, dtype=torch.uint8, device=rowptr.device)
    with torch.cuda.device(rowptr.device):
        _lib.check(lib.cr_spmm_plan(_ptr(rowptr), n_rows, nnz, d, _ptr(plan), nbytes, _stream(rowptr.device)), "cr_spmm_plan")
    return plan


def spmm(rowptr, col, val, X, Y=None, acc=None, acc_beta: float = 1.0, acc_div: float = 1.0, plan=None, acc_in=None):
    """y = A.X over the local row block; Y = y and/or acc = (acc_beta*acc + y)/acc_div (see the header)."""
    lib = _lib.load()
    rowptr = _req(rowptr, torch.int64, "rowptr")
    col = _req(col, torch.int32, "col")
    val = _req(val, torch.float32, "val", optional=True)
    X = _req(X, torch.float32, "X")
    Y = _req(Y, torch.float32, "Y", optional=True)
    acc = _req(acc, torch.float32, "acc", optional=True)
    acc_in = _req(acc_in, torch.float32, "acc_in", optional=True)
    dev = _same_device(rowptr, col, val, X, Y, acc, acc_in, plan)
    n_rows, nnz, d = rowptr.numel() - 1, col.numel(), X.shape[1]
    if val is not None and val.numel() != nnz:
        raise ValueError("val and col differ in length")
    for t, name in ((Y, "Y"), (acc, "acc"), (acc_in, "acc_in")):
        if t is not None and tuple(t.shape) != (n_rows, d):
            raise ValueError(f"{name} must be ({n_rows}, {d}), got {tuple(t.shape)}")
    if n_rows == 0:
        return Y if Y is not None else acc
    with torch.cuda.device(dev):
        rc = lib.cr_spmm_csr_f32(_ptr(rowptr), _ptr(col), _ptr(val), n_rows, nnz, _ptr(X), d, _ptr(Y), _ptr(acc_in), _ptr(acc),
                                 float(acc_beta), float(acc_div), _ptr(plan), 0 if plan is None else plan.numel(), _stream(dev))
    _lib.check(rc, "cr_spmm_csr_f32")
    return Y if Y is not None else acc


def spmm_bcast(rowptr, col, val, X, peer_tables_dev: int, n_peers: int, peer_row_offset: int, acc=None, acc_in=None,
               acc_beta: float = 1.0, acc_div: float = 1.0, plan=None, bcast_acc: bool = False, peer_row_split: Optional[int] = None,
               peer_row_offset_hi: int = 0, peer_need=None, multicast_ptr: int = 0):
    """SpMM on the local row block whose finished rows are stored straight into every peer's gather table
    (``peer_tables_dev`` = device address of an array of ``n_peers`` table pointers, e.g. symmetric-memory
    ``buffer_ptrs_dev``).  Local row r lands in destination row ``peer_row_offset + r`` (r < ``peer_row_split``) or
    ``peer_row_offset_hi + r``; by default there is one destination range.  ``peer_need`` (uint8 per local row, bit p =
    GPU p wants the row) restricts the stores to the GPUs that read the row; ``multicast_ptr`` (NVLS multicast address of the
    same tables, 0 = none) lets rows wanted everywhere go out as one switch-replicated store.  The caller barriers the GPUs
    afterwards."""
    lib = _lib.load()
    rowptr = _req(rowptr, torch.int64, "rowptr"); col = _req(col, torch.int32, "col")
    val = _req(val, torch.float32, "val", optional=True); X = _req(X, torch.float32, "X")
    acc = _req(acc, torch.float32, "acc", optional=True); acc_in = _req(acc_in, torch.float32, "acc_in", optional=True)
    peer_need = _req(peer_need, torch.uint8, "peer_need", optional=True)
    dev = _same_device(rowptr, col, val, X, acc, acc_in, plan, peer_need)
    n_rows, nnz, d = rowptr.numel() - 1, col.numel(), X.shape[1]
    if n_rows == 0:
        return acc
    if peer_need is not None and peer_need.numel() < n_rows:
        raise ValueError("peer_need must have one byte per local row")
    with torch.cuda.device(dev):
        rc = lib.cr_spmm_csr_bcast_f32(_ptr(rowptr), _ptr(col), _ptr(val), n_rows, nnz, _ptr(X), d, ctypes.c_void_p(peer_tables_dev),
                                       n_peers, peer_row_offset, n_rows if peer_row_split is None else peer_row_split,
                                       peer_row_offset_hi, int(bool(bcast_acc)), _ptr(peer_need), ctypes.c_void_p(multicast_ptr or None), _ptr(acc_in), _ptr(acc), float(acc_beta),
                                       float(acc_div), _ptr(plan), 0 if plan is None else plan.numel(), _stream(dev))
    _lib.check(rc, "cr_spmm_csr_bcast_f32")
    return acc


# ------------------------------------------------------------------------------------------------ K1
def score_topk(user_tab, item_tab, K: int, *, user_ids=None, item_gids=None, item_id_base: int = 0, mask_rowptr=None,
               mask_col=None, item_flags=None, flag_exclude: int = 0, precision: int = SCORE_EXACT_F32,
               out_score=None, out_id=None, workspace=None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Fused score -> mask -> top-K.  Returns (scores [n_q,K] fp32, ids [n_q,K] int32, n_refined int32[1])."""
    lib = _lib.load()
    user_tab = _req(user_tab, torch.float32, "user_tab")
    item_tab = _req(item_tab, torch.float32, "item_tab")
    user_ids = _req(user_ids, torch.int32, "user_ids", optional=True)
    item_gids = _req(item_gids, torch.int32, "item_gids", optional=True)
    mask_rowptr = _req(mask_rowptr, torch.int64, "mask_rowptr", optional=True)
    mask_col = _req(mask_col, torch.int32, "mask_col", optional=True)
    item_flags = _req(item_flags, torch.uint8, "item_flags", optional=True)
    dev = _same_device(user_tab, item_tab, user_ids, item_gids, mask_rowptr, mask_col, item_flags)
    if user_tab.dim() != 2 or item_tab.dim() != 2 or user_tab.shape[1] != item_tab.shape[1]:
        raise ValueError(f"user_tab {tuple(user_tab.shape)} and item_tab {tuple(item_tab.shape)} must share d")
    d, n_items = item_tab.shape[1], item_tab.shape[0]
    n_q = user_tab.shape[0] if user_ids is None else user_ids.numel()
    if item_gids is not None and item_gids.numel() != n_items:
        raise ValueError("item_gids must have one entry per item_tab row")
    if mask_rowptr is not None and mask_rowptr.numel() != n_q + 1:
        raise ValueError(f"mask_rowptr must have n_q+1={n_q + 1} entries")
    if mask_rowptr is not None and (mask_col is None or mask_col.numel() == 0):
        mask_rowptr = mask_col = None          # no train interactions at all (e.g. cold users): nothing to mask
    if out_score is None:
        out_score = torch.empty((n_q, K), dtype=torch.float32, device=dev)
    if out_id is None:
        out_id = torch.empty((n_q, K), dtype=torch.int32, device=dev)
    _req(out_score, torch.float32, "out_score"); _req(out_id, torch.int32, "out_id")
    n_ref = torch.zeros(1, dtype=torch.int32, device=dev)
    need = lib.cr_score_topk_workspace_bytes(n_q, n_items, d, K, precision)
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty(max(need, 1), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = lib.cr_score_topk_f32(_ptr(user_tab), _ptr(user_ids), n_q, _ptr(item_tab), _ptr(item_gids), item_id_base, n_items,
                                   d, _ptr(mask_rowptr), _ptr(mask_col), _ptr(item_flags), flag_exclude, K, _ptr(out_score),
                                   _ptr(out_id), _ptr(n_ref), precision, _ptr(workspace), workspace.numel(), _stream(dev))
    _lib.check(rc, "cr_score_topk_f32")
    return out_score, out_id, n_ref


def debug_tc_tile(user_tab, item_tab, K: int = 20):
    """Diagnostic: TF32-checked scoring (d=64, no masks) + raw tensor-core scores of the first 256x96 block."""
    lib = _lib.load()
    user_tab = _req(user_tab, torch.float32, "user_tab"); item_tab = _req(item_tab, torch.float32, "item_tab")
    dev = _same_device(user_tab, item_tab)
    n_q, n_items = user_tab.shape[0], item_tab.shape[0]
    out_s = torch.empty((n_q, K), dtype=torch.float32, device=dev)
    out_i = torch.empty((n_q, K), dtype=torch.int32, device=dev)
    dbg = torch.zeros((256, 96), dtype=torch.float32, device=dev)
    need = lib.cr_score_topk_workspace_bytes(n_q, n_items, 64, K, SCORE_TF32_CHECKED)
    ws = torch.empty(max(need, 1), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = lib.cr_debug_tc_tile(_ptr(user_tab), n_q, _ptr(item_tab), n_items, K, _ptr(out_s), _ptr(out_i), _ptr(dbg), _ptr(ws),
                                  ws.numel(), _stream(dev))
    _lib.check(rc, "cr_debug_tc_tile")
    return out_s, out_i, dbg


def topk_merge(in_score: torch.Tensor, in_id: torch.Tensor, out_score=None, out_id=None):
    """Merge [G, n_q, K] candidate lists into [n_q, K] by (score desc, id asc)."""
    lib = _lib.load()
    in_score = _req(in_score, torch.float32, "in_score")
    in_id = _req(in_id, torch.int32, "in_id")
    if in_score.dim() != 3 or in_score.shape != in_id.shape:
        raise ValueError("in_score/in_id must both be [G, n_q, K]")
    dev = _same_device(in_score, in_id)
    G, n_q, K = in_score.shape
    if out_score is None:
        out_score = torch.empty((n_q, K), dtype=torch.float32, device=dev)
    if out_id is None:
        out_id = torch.empty((n_q, K), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = lib.cr_topk_merge(_ptr(in_score), _ptr(in_id), G, n_q, K, _ptr(out_score), _ptr(out_id), _stream(dev))
    _lib.check(rc, "cr_topk_merge")
    return out_score, out_id


def fill_masked(out_score, out_id, n_items_total: int, item_flags=None, flag_exclude: int = 0, mask_rowptr=None, mask_col=None):
    """Complete short lists (trailing id -1) with masked ids at CR_MASK_SCORE, in place."""
    lib = _lib.load()
    _req(out_score, torch.float32, "out_score"); _req(out_id, torch.int32, "out_id")
    item_flags = _req(item_flags, torch.uint8, "item_flags", optional=True)
    mask_rowptr = _req(mask_rowptr, torch.int64, "mask_rowptr", optional=True)
    mask_col = _req(mask_col, torch.int32, "mask_col", optional=True)
    dev = _same_device(out_score, out_id, item_flags, mask_rowptr, mask_col)
    n_q, K = out_id.shape
    with torch.cuda.device(dev):
        rc = lib.cr_fill_masked(_ptr(out_score), _ptr(out_id), n_q, K, n_items_total, _ptr(item_flags), flag_exclude,
                                _ptr(mask_rowptr), _ptr(mask_col), _stream(dev))
    _lib.check(rc, "cr_fill_masked")
    return out_score, out_id


def gather_rows(src: torch.Tensor, ids: torch.Tensor, out=None) -> torch.Tensor:
    lib = _lib.load()
    src = _req(src, torch.float32, "src")
    ids = _req(ids, torch.int32, "ids")
    dev = _same_device(src, ids)
    n, d = ids.numel(), src.shape[1]
    if out is None:
        out = torch.empty((n, d), dtype=torch.float32, device=dev)
    if n == 0:
        return out
    with torch.cuda.device(dev):
        _lib.check(lib.cr_gather_rows_f32(_ptr(src), _ptr(ids), n, d, _ptr(out), _stream(dev)), "cr_gather_rows_f32")
    return out


def copy_rows(src: torch.Tensor, dst: torch.Tensor, src_ids=None, dst_ids=None) -> torch.Tensor:
    """dst[dst_ids[j]] = src[src_ids[j]] (either id list may be None = j); ``dst_ids`` must not repeat.  In place on dst."""
    lib = _lib.load()
    src = _req(src, torch.float32, "src"); dst = _req(dst, torch.float32, "dst")
    src_ids = _req(src_ids, torch.int32, "src_ids", optional=True); dst_ids = _req(dst_ids, torch.int32, "dst_ids", optional=True)
    dev = _same_device(src, dst, src_ids, dst_ids)
    if src.dim() != 2 or dst.dim() != 2 or src.shape[1] != dst.shape[1]:
        raise ValueError(f"src {tuple(src.shape)} and dst {tuple(dst.shape)} must be 2-D tables of the same width")
    counts = {t.numel() for t in (src_ids, dst_ids) if t is not None}
    if len(counts) > 1:
        raise ValueError("src_ids and dst_ids differ in length")
    n = counts.pop() if counts else min(src.shape[0], dst.shape[0])
    if (src_ids is None and n > src.shape[0]) or (dst_ids is None and n > dst.shape[0]):
        raise ValueError("more rows to copy than the table without an id list holds")
    if n == 0:
        return dst
    with torch.cuda.device(dev):
        _lib.check(lib.cr_copy_rows_f32(_ptr(src), _ptr(src_ids), _ptr(dst_ids), n, src.shape[1], _ptr(dst), _stream(dev)),
                   "cr_copy_rows_f32")
    return dst


# ------------------------------------------------------------------------------------------------ K2
_DCG_TABLES = {}


def dcg_tables(K: int):
    """1/log(n+2,2) and its running sums, computed with the reference's own expression and addition
    order (util/evaluator.py:104-109) so device-side DCG/IDCG match it bit for bit."""
    inv = [1.0 / math.log(n + 2, 2) for n in range(K)]
    pre, s = [0.0], 0
    for v in inv:
        s += v
        pre.append(float(s))
    return inv, pre


def rank_metrics(topk_id: torch.Tensor, gt_rowptr: torch.Tensor, gt_col: torch.Tensor, Ns: Sequence[int],
                 per_query: bool = False):
    """Device reduction of Hit/Precision/Recall/NDCG partials.  Returns (sums [nN,6] fp64 on device,
    hits [nN,n_q] int32 | None, dcg [nN,n_q] fp64 | None)."""
    lib = _lib.load()
    topk_id = _req(topk_id, torch.int32, "topk_id")
    gt_rowptr = _req(gt_rowptr, torch.int64, "gt_rowptr")
    gt_col = _req(gt_col, torch.int32, "gt_col")
    dev = _same_device(topk_id, gt_rowptr, gt_col)
    n_q, K = topk_id.shape
    if gt_rowptr.numel() != n_q + 1:
        raise ValueError("gt_rowptr must have n_q+1 entries")
    nN = len(Ns)
    tab = _DCG_TABLES.get((K, dev))
    if tab is None:                  # constants of (K): uploaded once per device, not once per evaluation (a synchronous pageable copy)
        inv, pre = dcg_tables(K)
        tab = _DCG_TABLES[(K, dev)] = torch.tensor(inv + pre, dtype=torch.float64, device=dev)
    sums = torch.empty((nN, 6), dtype=torch.float64, device=dev)
    hits = torch.empty((nN, n_q), dtype=torch.int32, device=dev) if per_query else None
    dcg = torch.empty((nN, n_q), dtype=torch.float64, device=dev) if per_query else None
    ws = torch.empty(lib.cr_rank_metrics_workspace_bytes(n_q, nN), dtype=torch.uint8, device=dev)
    ns_arr = (ctypes.c_int32 * nN)(*[int(n) for n in Ns])
    with torch.cuda.device(dev):
        rc = lib.cr_rank_metrics(_ptr(topk_id), n_q, K, _ptr(gt_rowptr), _ptr(gt_col), ns_arr, nN, _ptr(tab),
                                 ctypes.c_void_p(tab.data_ptr() + 8 * K), _ptr(hits), _ptr(dcg), _ptr(sums), _ptr(ws),
                                 ws.numel(), _stream(dev))
    _lib.check(rc, "cr_rank_metrics")
    return sums, hits, dcg


# ------------------------------------------------------------------------------------------------ K4
def linear_act(X1, W, bias=None, *, X2=None, xrow=None, scale=None, shift=None, act=None, out=None, yrow=None):
    """out[yrow] = act(([X1|X2][xrow] . W^T + bias) * scale + shift); W in nn.Linear layout [n_out, d1+d2]."""
    lib = _lib.load()
    X1 = _req(X1, torch.float32, "X1", contiguous=False)
    X2 = _req(X2, torch.float32, "X2", optional=True, contiguous=False)
    W = _req(W, torch.float32, "W")
    bias = _req(bias, torch.float32, "bias", optional=True)
    scale = _req(scale, torch.float32, "scale", optional=True)
    shift = _req(shift, torch.float32, "shift", optional=True)
    xrow = _req(xrow, torch.int32, "xrow", optional=True)
    yrow = _req(yrow, torch.int32, "yrow", optional=True)
    dev = _same_device(X1, X2, W, bias, scale, shift, xrow, yrow, out)
    for t, name in ((X1, "X1"), (X2, "X2")):
        if t is not None and (t.dim() != 2 or t.stride(1) != 1):
            raise ValueError(f"{name} must be 2-D with unit inner stride")
    d1, d2 = X1.shape[1], (0 if X2 is None else X2.shape[1])
    n_out = W.shape[0]
    if W.shape[1] != d1 + d2:
        raise ValueError(f"W is {tuple(W.shape)}, inputs give k={d1 + d2}")
    n_rows = xrow.numel() if xrow is not None else X1.shape[0]
    if out is None:
        if yrow is not None:
            raise ValueError("a scatter (yrow) needs an existing `out` table")
        out = torch.empty((n_rows, n_out), dtype=torch.float32, device=dev)
    _req(out, torch.float32, "out", contiguous=False)
    if act not in _ACTS:
        raise ValueError(f"unknown activation {act!r}")
    with torch.cuda.device(dev):
        rc = lib.cr_linear_act_f32(_ptr(X1), X1.stride(0), d1, _ptr(X2), 0 if X2 is None else X2.stride(0), d2, _ptr(xrow),
                                   n_rows, _ptr(W), _ptr(bias), _ptr(scale), _ptr(shift), n_out, _ACTS[act], _ptr(out),
                                   out.stride(0), _ptr(yrow), _stream(dev))
    _lib.check(rc, "cr_linear_act_f32")
    return out


class SplitTable:
    """A table as the tensor-core tower kernel consumes it: ``x == hi + lo`` exactly, ``hi`` representable in TF32.  Both
    halves are (rows, ld) fp32 with ld = width rounded up to a multiple of 4 (zero columns): a legal TMA row stride."""
    __slots__ = ("hi", "lo", "width")

    def __init__(self, hi: torch.Tensor, lo: torch.Tensor, width: int):
        self.hi, self.lo, self.width = hi, lo, int(width)

    @property
    def rows(self) -> int:
        return self.hi.shape[0]


def split_tf32(x: torch.Tensor) -> SplitTable:
    """hi = round-to-nearest-TF32(x), lo = x - hi for a (rows, cols) fp32 table with unit inner stride (``cr_split_tf32``).
    Do it once for constant tables (item content) and keep the result."""
    lib = _lib.load()
    x = _req(x, torch.float32, "x", contiguous=False)
    if x.dim() != 2 or x.stride(1) != 1:
        raise ValueError("x must be 2-D with unit inner stride")
    rows, cols = x.shape
    ld = (cols + 3) // 4 * 4
    hi = torch.empty((rows, ld), dtype=torch.float32, device=x.device)
    lo = torch.empty((rows, ld), dtype=torch.float32, device=x.device)
    if rows:
        with torch.cuda.device(x.device):
            rc = lib.cr_split_tf32(_ptr(x), max(x.stride(0), cols), rows, cols, _ptr(hi), _ptr(lo), ld, _stream(x.device))
        _lib.check(rc, "cr_split_tf32")
    return SplitTable(hi, lo, cols)


def tma_rows(x: torch.Tensor) -> torch.Tensor:
    """``x`` as the tensor-core layer can read it directly: fp32, unit inner stride, 16-byte aligned rows (row stride a
    multiple of 4 floats).  Returns ``x`` itself when it already qualifies, else a padded copy (a view of width ``x.shape[1]``
    into a (rows, ld) buffer) — do it once for a constant table such as the item content (2,738 columns -> stride 2,740)."""
    x = _req(x, torch.float32, "x", contiguous=False)
    if x.dim() != 2:
        raise ValueError("x must be 2-D")
    if x.stride(1) == 1 and x.stride(0) % 4 == 0 and x.data_ptr() % 16 == 0:
        return x
    rows, cols = x.shape
    ld = (cols + 3) // 4 * 4
    buf = torch.zeros((rows, ld), dtype=torch.float32, device=x.device)
    buf[:, :cols].copy_(x)
    return buf[:, :cols]


def linear_act_tc(X1, W: SplitTable, bias=None, *, X2=None, scale=None, shift=None, act=None,
                  out=None, yrow=None, want_split: bool = False, want_plain: bool = True):
    """``linear_act`` on the tensor cores at fp32 accuracy (``cr_linear_act_tc_f32``: tcgen05 kind::tf32, hi.hi + lo.hi + hi.lo,
    K blocks drained into fp32 running sums).  W is the ``SplitTable`` of the nn.Linear weight [n_out, k1 + k2].  X1 / X2 are
    ``SplitTable``s (what towers.py passes: fastest) or plain fp32 tensors, split into hi + lo INSIDE the kernel (one HBM read,
    bit-identical, measured ~15 % slower on wide-K layers).  Returns
    (out | None, SplitTable of the output | None); rows are contiguous (no gather), ``yrow`` scatters the rows of ``out``
    (GAR.py:44-46)."""
    lib = _lib.load()
    raw = not isinstance(X1, SplitTable)
    if X2 is not None and isinstance(X2, SplitTable) == raw:       # mixed: bring both to the split form
        X1, X2, raw = (X1 if isinstance(X1, SplitTable) else split_tf32(X1)), (X2 if isinstance(X2, SplitTable) else split_tf32(X2)), False
    if raw:
        X1 = tma_rows(X1)
        X2 = None if X2 is None else tma_rows(X2)
        x1h, x1l, ld1, d1, n_rows = X1, None, X1.stride(0) if X1.shape[0] > 1 else max(X1.stride(0), (X1.shape[1] + 3) // 4 * 4), X1.shape[1], X1.shape[0]
        x2h, x2l, ld2, d2 = (None, None, 0, 0) if X2 is None else (X2, None, X2.stride(0) if X2.shape[0] > 1 else max(X2.stride(0), (X2.shape[1] + 3) // 4 * 4), X2.shape[1])
        if X2 is not None and X2.shape[0] != n_rows:
            raise ValueError("X1 and X2 differ in rows")
        tabs = [t for t in (x1h, x2h) if t is not None]
    else:
        x1h, x1l, ld1, d1, n_rows = X1.hi, X1.lo, X1.hi.stride(0), X1.width, X1.rows
        x2h, x2l, ld2, d2 = (None, None, 0, 0) if X2 is None else (X2.hi, X2.lo, X2.hi.stride(0), X2.width)
        if X2 is not None and X2.rows != n_rows:
            raise ValueError("X1 and X2 differ in rows")
        tabs = [t for t in (x1h, x1l, x2h, x2l) if t is not None]
        for t in tabs:
            _req(t, torch.float32, "split table")
    for t in (W.hi, W.lo):
        _req(t, torch.float32, "W split table")
    bias = _req(bias, torch.float32, "bias", optional=True)
    scale = _req(scale, torch.float32, "scale", optional=True)
    shift = _req(shift, torch.float32, "shift", optional=True)
    yrow = _req(yrow, torch.int32, "yrow", optional=True)
    dev = _same_device(*tabs, W.hi, W.lo, bias, scale, shift, yrow, out)
    n_out = W.rows
    if W.width != d1 + d2:
        raise ValueError(f"W has {W.width} columns, inputs give k={d1 + d2}")
    if act not in _ACTS:
        raise ValueError(f"unknown activation {act!r}")
    if out is None and want_plain:
        if yrow is not None:
            raise ValueError("a scatter (yrow) needs an existing `out` table")
        out = torch.empty((n_rows, n_out), dtype=torch.float32, device=dev)
    if out is not None:
        _req(out, torch.float32, "out", contiguous=False)
    sp = None
    if want_split:
        ldh = (n_out + 3) // 4 * 4
        sp = SplitTable(torch.empty((n_rows, ldh), dtype=torch.float32, device=dev), torch.empty((n_rows, ldh), dtype=torch.float32, device=dev), n_out)
    if n_rows == 0:
        return out, sp
    with torch.cuda.device(dev):
        rc = lib.cr_linear_act_tc_f32(_ptr(x1h), _ptr(x1l), ld1, d1, _ptr(x2h), _ptr(x2l), ld2, d2, n_rows,
                                      _ptr(W.hi), _ptr(W.lo), W.hi.stride(0), _ptr(bias), _ptr(scale), _ptr(shift), n_out, _ACTS[act],
                                      _ptr(out), 0 if out is None else out.stride(0), _ptr(yrow), _ptr(None if sp is None else sp.hi),
                                      _ptr(None if sp is None else sp.lo), 0 if sp is None else sp.hi.stride(0), _stream(dev))
    _lib.check(rc, "cr_linear_act_tc_f32")
    return out, sp


def topk_rows(S: torch.Tensor, K: int, exclude_col=None, col_id_base: int = 0):
    """Row-wise top-K of a dense score block: (scores [n_rows, K], ids [n_rows, K] int32), (score desc, id asc)."""
    lib = _lib.load()
    S = _req(S, torch.float32, "S", contiguous=False)
    if S.dim() != 2 or S.stride(1) != 1:
        raise ValueError("S must be 2-D with unit inner stride")
    exclude_col = _req(exclude_col, torch.int32, "exclude_col", optional=True)
    dev = _same_device(S, exclude_col)
    n_rows, n_cols = S.shape
    out_s = torch.empty((n_rows, K), dtype=torch.float32, device=dev)
    out_i = torch.empty((n_rows, K), dtype=torch.int32, device=dev)
    if n_rows:
        with torch.cuda.device(dev):
            rc = lib.cr_topk_rows_f32(_ptr(S), n_rows, n_cols, max(S.stride(0), n_cols), K, _ptr(exclude_col), int(col_id_base), _ptr(out_s),
                                      _ptr(out_i), _stream(dev))
        _lib.check(rc, "cr_topk_rows_f32")
    return out_s, out_i


def bn_fold(gamma, beta, mean, var, eps: float):
    """Eval-mode BatchNorm1d as (scale, shift)."""
    lib = _lib.load()
    mean = _req(mean, torch.float32, "running_mean"); var = _req(var, torch.float32, "running_var")
    gamma = _req(gamma, torch.float32, "weight", optional=True); beta = _req(beta, torch.float32, "bias", optional=True)
    dev = _same_device(mean, var, gamma, beta)
    n = mean.numel()
    scale = torch.empty(n, dtype=torch.float32, device=dev)
    shift = torch.empty(n, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = lib.cr_bn_fold_f32(_ptr(gamma), _ptr(beta), _ptr(mean), _ptr(var), float(eps), n, _ptr(scale), _ptr(shift), _stream(dev))
    _lib.check(rc, "cr_bn_fold_f32")
    return scale, shift


def heater_blend(gate, expert, Vin, keep: float, one_minus_keep: float):
    lib = _lib.load()
    gate = _req(gate, torch.float32, "gate"); expert = _req(expert, torch.float32, "expert"); Vin = _req(Vin, torch.float32, "Vin")
    dev = _same_device(gate, expert, Vin)
    n, d = expert.shape
    if Vin.shape != expert.shape or gate.shape[0] != n:
        raise ValueError("gate/expert/Vin row counts differ")
    out = torch.empty_like(expert)
    with torch.cuda.device(dev):
        rc = lib.cr_heater_blend_f32(_ptr(gate), gate.shape[1], _ptr(expert), _ptr(Vin), float(keep), float(one_minus_keep), n, d,
                                     _ptr(out), _stream(dev))
    _lib.check(rc, "cr_heater_blend_f32")
    return out


# ------------------------------------------------------------------------------------------------ K5 / K6
def bpr_fwd_bwd(user_emb, item_emb, u_idx, i_idx, j_idx, reg: float, grad_user, grad_item, loss=None, workspace=None):
    """BPR + L2 loss of one batch and its gradients w.r.t. the gathered rows, accumulated into the dense
    gradient tables (which the caller zeroed).  Returns loss fp32[4] on device: total, bpr, reg, 0."""
    lib = _lib.load()
    user_emb = _req(user_emb, torch.float32, "user_emb"); item_emb = _req(item_emb, torch.float32, "item_emb")
    u_idx = _req(u_idx, torch.int32, "u_idx"); i_idx = _req(i_idx, torch.int32, "i_idx"); j_idx = _req(j_idx, torch.int32, "j_idx")
    grad_user = _req(grad_user, torch.float32, "grad_user"); grad_item = _req(grad_item, torch.float32, "grad_item")
    dev = _same_device(user_emb, item_emb, u_idx, i_idx, j_idx, grad_user, grad_item)
    if user_emb.dim() != 2 or item_emb.dim() != 2 or user_emb.shape[1] != item_emb.shape[1]:
        raise ValueError("user_emb and item_emb must be 2-D with the same width")
    if grad_user.shape != user_emb.shape or grad_item.shape != item_emb.shape:
        raise ValueError("gradient tables must have the shapes of the embedding tables")
    B = u_idx.numel()
    if i_idx.numel() != B or j_idx.numel() != B:
        raise ValueError("u_idx / i_idx / j_idx differ in length")
    if loss is None:
        loss = torch.zeros(4, dtype=torch.float32, device=dev)
    _req(loss, torch.float32, "loss")
    need = lib.cr_bpr_workspace_bytes(B)
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty(max(need, 1), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = lib.cr_bpr_fwd_bwd_f32(_ptr(user_emb), _ptr(item_emb), user_emb.shape[1], _ptr(u_idx), _ptr(i_idx), _ptr(j_idx), B,
                                    float(reg), _ptr(loss), _ptr(grad_user), _ptr(grad_item), _ptr(workspace), workspace.numel(),
                                    _stream(dev))
    _lib.check(rc, "cr_bpr_fwd_bwd_f32")
    return loss


def adam_scalars(step: int, lr: float, betas=(0.9, 0.999)):
    """(lr / (1 - beta1^t), sqrt(1 - beta2^t)) as the library computes them, for ``dev_scalars`` of adam_step."""
    out = (ctypes.c_float * 2)()
    _lib.check(_lib.load().cr_adam_scalars(float(lr), float(betas[0]), float(betas[1]), int(step), out), "cr_adam_scalars")
    return float(out[0]), float(out[1])


def adam_step(param, grad, exp_avg, exp_avg_sq, step: int, lr: float, betas=(0.9, 0.999), eps: float = 1e-8, grad_scale: float = 1.0,
              dev_scalars=None):
    """In-place torch.optim.Adam update of one fp32 tensor (state tensors updated in place as well).  With ``dev_scalars``
    (device fp32[2] holding ``adam_scalars(step, ...)``) the step-dependent factors are read on the device (graph replay)."""
    lib = _lib.load()
    for t, name in ((param, "param"), (grad, "grad"), (exp_avg, "exp_avg"), (exp_avg_sq, "exp_avg_sq")):
        _req(t, torch.float32, name)
        if t.numel() != param.numel():
            raise ValueError(f"{name} has {t.numel()} elements, param has {param.numel()}")
    dev_scalars = _req(dev_scalars, torch.float32, "dev_scalars", optional=True)
    dev = _same_device(param, grad, exp_avg, exp_avg_sq, dev_scalars)
    with torch.cuda.device(dev):
        rc = lib.cr_adam_step_f32(_ptr(param), _ptr(grad), _ptr(exp_avg), _ptr(exp_avg_sq), param.numel(), float(lr), float(betas[0]),
                                  float(betas[1]), float(eps), int(step), float(grad_scale), _ptr(dev_scalars), _stream(dev))
    _lib.check(rc, "cr_adam_step_f32")
    return param


def sample_pairwise(pair_user, pair_item, train_rowptr, train_col, n_items: int, seed: int, epoch: int, begin: int, count: int,
                    out=None, n_exhausted=None):
    """(user, positive, negative) triples of positions [begin, begin+count) of epoch `epoch`'s shuffled pair list."""
    lib = _lib.load()
    pair_user = _req(pair_user, torch.int32, "pair_user"); pair_item = _req(pair_item, torch.int32, "pair_item")
    train_rowptr = _req(train_rowptr, torch.int64, "train_rowptr"); train_col = _req(train_col, torch.int32, "train_col")
    n_exhausted = _req(n_exhausted, torch.int32, "n_exhausted", optional=True)
    dev = _same_device(pair_user, pair_item, train_rowptr, train_col, n_exhausted)
    n_pairs = pair_user.numel()
    if pair_item.numel() != n_pairs:
        raise ValueError("pair_user / pair_item differ in length")
    if out is None:
        out = torch.empty((3, count), dtype=torch.int32, device=dev)
    _req(out, torch.int32, "out")
    if out.shape[0] != 3 or out.shape[1] < count:
        raise ValueError("out must be int32 [3, >=count]")
    with torch.cuda.device(dev):
        rc = lib.cr_sample_pairwise(_ptr(pair_user), _ptr(pair_item), n_pairs, _ptr(train_rowptr), _ptr(train_col), int(n_items),
                                    int(seed) & (2**64 - 1), int(epoch) & (2**64 - 1), int(begin), int(count), _ptr(out[0]), _ptr(out[1]),
                                    _ptr(out[2]), _ptr(n_exhausted), _stream(dev))
    _lib.check(rc, "cr_sample_pairwise")
    return out[0, :count], out[1, :count], out[2, :count]
