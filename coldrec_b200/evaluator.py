"""Ranking metrics with the reference's interface, reduced on the device.

``ranking_evaluation(origin, res, N)`` keeps the signature and return format of
util/evaluator.py:153-187 — ``measure`` = ['Top 10\\n', 'Hit Ratio:v\\n', 'Precision:v\\n',
'Recall:v\\n', 'NDCG:v\\n', ...] and ``performance[i] = [hr, prec, recall, ndcg]``, every value
``round(., 5)`` — but the per-user work (set intersections, DCG loops; 24 us/user in the reference)
runs in ``cr_rank_metrics`` on the top-K id tensor.  ``res`` may be the ``RecList`` returned by the
fused ``_evaluate`` (ids stay on the device) or a plain reference-style dict.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import ops
from .scoring import EvalPlan, _csr_from_lists


class RecList(dict):
    """``{raw_user: [(raw_item, score), ...]}`` whose Python form is only built when somebody reads
    it (model/BaseRecommender.py:183-187 builds it eagerly: K dict lookups per user).  The device
    tensors stay attached so metrics never leave the GPU."""

    def __init__(self, plan: EvalPlan, scores: torch.Tensor, ids: torch.Tensor, id2item: Optional[Dict[int, int]]):
        super().__init__()
        self.plan, self.scores, self.ids, self._id2item = plan, scores, ids, id2item
        self._built = False

    def _build(self):
        if self._built:
            return
        self._built = True
        ids, scores = self.ids.cpu().numpy(), self.scores.cpu().numpy()
        users = self.plan.users if self.plan.users is not None else list(range(ids.shape[0]))
        m = self._id2item
        for u, row_i, row_s in zip(users, ids, scores):
            names = [m[int(i)] for i in row_i] if m is not None else [int(i) for i in row_i]
            dict.__setitem__(self, u, list(zip(names, row_s)))

    def __getitem__(self, k):
        self._build(); return dict.__getitem__(self, k)

    def __iter__(self):
        self._build(); return dict.__iter__(self)

    def __len__(self):
        return self.plan.n_q

    def __contains__(self, k):
        self._build(); return dict.__contains__(self, k)

    def keys(self):
        self._build(); return dict.keys(self)

    def values(self):
        self._build(); return dict.values(self)

    def items(self):
        self._build(); return dict.items(self)

    def get(self, k, default=None):
        self._build(); return dict.get(self, k, default)


def metrics_from_sums(sums: np.ndarray, n_users: int, Ns: Sequence[int], rounded: bool = True) -> List[List[float]]:
    """Finish util/evaluator.py:18-32, 47-51, 54-63, 95-115 from the device partials
    [sum_hits, sum_gt, sum_recall, n_recall, sum_ndcg, n_ndcg] per cut-off."""
    out = []
    for (sh, sg, sr, nr, sn, nn), n in zip(sums.tolist(), Ns):
        hr = sh / sg if sg else 0.0
        prec = sh / (n_users * n) if (n_users and n) else 0.0
        recall = sr / nr if nr else 0.0
        ndcg = sn / nn if nn else 0.0
        vals = [hr, prec, recall, ndcg]
        out.append([round(v, 5) for v in vals] if rounded else vals)
    return out


def format_measure(perf: List[List[float]], Ns: Sequence[int]) -> List[str]:
    measure = []
    for n, (hr, prec, recall, ndcg) in zip(Ns, perf):
        measure.append('Top ' + str(n) + '\n')
        measure += ['Hit Ratio:' + str(hr) + '\n', 'Precision:' + str(prec) + '\n', 'Recall:' + str(recall) + '\n',
                    'NDCG:' + str(ndcg) + '\n']
    return measure


def device_metrics(topk_ids: torch.Tensor, gt_rowptr: torch.Tensor, gt_col: torch.Tensor, Ns: Sequence[int],
                   rounded: bool = True) -> List[List[float]]:
    sums, _, _ = ops.rank_metrics(topk_ids, gt_rowptr, gt_col, Ns)
    return metrics_from_sums(sums.cpu().numpy(), topk_ids.shape[0], Ns, rounded)


def _plan_matches(origin: Dict, res: RecList) -> bool:
    return res.plan.users is not None and len(origin) == len(res.plan.users) and getattr(res, "_gt_id", None) == id(origin)


def ranking_evaluation(origin: Dict, res, N: Sequence[int], item_map: Optional[Dict] = None, device=None):
    """Drop-in for util/evaluator.py:153-187.  ``origin`` = {user: {item: rating}}, ``res`` = RecList or
    {user: [(item, score), ...]}; ``item_map`` (raw item -> dense id) is needed only for plain dicts."""
    if len(origin) != len(res):                                  # :161-164
        print(f"ground-truth set size: {len(origin)}, predicted set size: {len(res)}")
        print('The Lengths of ground-truth set and predicted set do not match!')
        raise SystemExit(-1)
    if isinstance(res, RecList) and _plan_matches(origin, res):
        ids, rowptr, col = res.ids, res.plan.gt_rowptr, res.plan.gt_col
    else:
        # reference-style dict: upload ids once, then the same device reduction
        users = list(origin.keys())
        if item_map is None:
            universe = {}
            for u in users:
                for it in origin[u]:
                    universe.setdefault(it, len(universe))
                for it, _ in res[u]:
                    universe.setdefault(it, len(universe))
            item_map = universe
        K = max(N)
        rows = []
        for u in users:
            r = [item_map[it] for it, _ in res[u][:K]]
            rows.append(r + [-1] * (K - len(r)))
        if device is None:
            device = res.ids.device if isinstance(res, RecList) else torch.device("cuda", torch.cuda.current_device())
        ids = torch.tensor(rows, dtype=torch.int32, device=device).reshape(len(users), K)
        rp, c = _csr_from_lists([np.sort(np.fromiter((item_map[i] for i in origin[u]), dtype=np.int64, count=len(origin[u])))
                                 for u in users])
        rowptr, col = torch.from_numpy(rp).to(device), torch.from_numpy(c).to(device)
    perf = device_metrics(ids, rowptr, col, list(N))
    return format_measure(perf, N), perf
