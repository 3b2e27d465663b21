"""Array-based data / graph build path.

``ArrayDataBuilder`` takes the constructor arguments of ``ColdStartDataBuilder`` (util/databuilder.py:7-10) and produces
the same id tables, index lists and matrices, but every O(E) step is a numpy array operation instead of a Python loop
over interactions:

  * ``generate_set`` (:90-216): dense ids in first-seen order over training, warm-valid, warm-test, cold-valid,
    cold-test, overall-valid, overall-test  ->  one ``np.unique(return_index)`` over the concatenated columns;
  * the adjacency list comprehensions (:225-226, :270-273)  ->  id lookups through ``np.searchsorted``;
  * the per-user mask tensors of ``_get_eval_cache`` (model/BaseRecommender.py:115-128)  ->  ``eval_plan`` builds the
    scorer's CSR (train mask + ground truth) for a whole split with one sort;
  * ``TorchGraphInterface.convert_sparse_mat_to_tensor`` (:953-962, int64 COO, 20 B/nnz)  ->  ``graph(device)`` hands the
    int32 CSR (8 B/nnz) to the SpMM kernel, and ``sampler(device)`` the pair list to the device sampler.

The dict-of-dict views the reference exposes (``training_set_u``, ``warm_test_set`` ...) are still available — built on
first access — because ``ranking_evaluation(origin, res, N)`` and model code consume them; the hot path does not.
On-disk formats stay the reference's (``DataLoader.load_data_set`` triples, ``info_dict.pkl`` lists, ``.npy`` content).
"""
from __future__ import annotations

from collections import defaultdict
from typing import Dict, Optional, Sequence

import numpy as np
import scipy.sparse as sp

SPLITS = ("training", "warm_valid", "warm_test", "cold_valid", "cold_test", "overall_valid", "overall_test")   # :90-216 order


def _pairs(rows) -> np.ndarray:
    """(n, 2) int64 array of (user, item) from a list of [user, item, rating] triples or an (n, >=2) array."""
    if isinstance(rows, np.ndarray):
        a = rows
    else:
        a = np.asarray([(r[0], r[1]) for r in rows], dtype=np.int64) if len(rows) else np.zeros((0, 2), dtype=np.int64)
    if a.ndim != 2 or a.shape[1] < 2:
        raise ValueError("interaction data must be (n, >=2): user, item[, rating]")
    return np.ascontiguousarray(a[:, :2]).astype(np.int64, copy=False)


def load_data_set(file) -> np.ndarray:
    """Vectorised ``DataLoader.load_data_set`` (util/loader.py:22-34): a header line, then ``user,item[,...]`` rows.
    Returns the (n, 2) int64 (user, item) array the builder takes directly (the reference builds a Python list of
    ``[user, item, 1.0]`` triples line by line — minutes and tens of GB at 10^8 interactions)."""
    import pandas as pd
    df = pd.read_csv(file, header=0, usecols=[0, 1], dtype=np.int64, engine="c")
    return np.ascontiguousarray(df.to_numpy(dtype=np.int64))


class _IdTable:
    """raw id -> dense id in first-seen order; vectorised lookups through a sorted copy."""

    def __init__(self, raw_stream: np.ndarray):
        uniq, first = np.unique(raw_stream, return_index=True)
        order = np.argsort(first, kind="stable")
        self.raw = uniq[order]                              # dense id -> raw id   (id2user / id2item)
        self._sorted_raw = uniq
        self._sorted_to_dense = np.empty(len(uniq), dtype=np.int64)
        self._sorted_to_dense[order] = np.arange(len(uniq))
        self._dict: Optional[Dict[int, int]] = None

    def __len__(self):
        return len(self.raw)

    def lookup(self, raw: np.ndarray, what: str) -> np.ndarray:
        raw = np.asarray(raw, dtype=np.int64)
        pos = np.searchsorted(self._sorted_raw, raw)
        pos_c = np.minimum(pos, max(len(self._sorted_raw) - 1, 0))
        ok = (len(self._sorted_raw) > 0) & (self._sorted_raw[pos_c] == raw) if len(raw) else np.zeros(0, bool)
        if len(raw) and not ok.all():
            raise Exception(f"{what} {raw[~ok][0]} not in current id table")        # util/databuilder.py:281,287,295,304
        return self._sorted_to_dense[pos_c] if len(raw) else np.zeros(0, dtype=np.int64)

    def as_dict(self) -> Dict[int, int]:
        if self._dict is None:
            self._dict = dict(zip(self.raw.tolist(), range(len(self.raw))))
        return self._dict


class ArrayDataBuilder:
    """Drop-in for ``ColdStartDataBuilder`` with array internals; same positional arguments (util/databuilder.py:7-10)."""

    def __init__(self, training_data, warm_valid_data, cold_valid_data, overall_valid_data, warm_test_data, cold_test_data,
                 overall_test_data, user_num, item_num, warm_user_idx, warm_item_idx, cold_user_idx, cold_item_idx,
                 user_content=None, item_content=None):
        self.user_num, self.item_num = int(user_num), int(item_num)
        self.training_data = training_data
        self.warm_valid_data, self.warm_test_data = warm_valid_data, warm_test_data
        self.cold_valid_data, self.cold_test_data = cold_valid_data, cold_test_data
        self.overall_valid_data, self.overall_test_data = overall_valid_data, overall_test_data
        self._pairs = {"training": _pairs(training_data), "warm_valid": _pairs(warm_valid_data), "warm_test": _pairs(warm_test_data),
                       "cold_valid": _pairs(cold_valid_data), "cold_test": _pairs(cold_test_data),
                       "overall_valid": _pairs(overall_valid_data), "overall_test": _pairs(overall_test_data)}
        stream = np.concatenate([self._pairs[k] for k in SPLITS], axis=0)
        self._users, self._items = _IdTable(stream[:, 0]), _IdTable(stream[:, 1])
        # dense-id pair arrays per split
        self._dense = {k: (self._users.lookup(v[:, 0], "user"), self._items.lookup(v[:, 1], "item")) for k, v in self._pairs.items()}
        self.train_user, self.train_item = self._dense["training"]

        self.source_user_content, self.source_item_content = user_content, item_content
        self.mapped_user_content = self.mapped_item_content = None
        if user_content is not None:            # :44-48, 58-65: row uid holds source row raw id
            n = max(self.user_num, user_content.shape[0], len(self._users))
            self.mapped_user_content = np.zeros((n, user_content.shape[1]), dtype=np.float64)
            self.mapped_user_content[:len(self._users)] = user_content[self._users.raw]
            self.user_content_dim = user_content.shape[-1]
        if item_content is not None:            # :49-54, 66-73
            n = max(self.item_num, item_content.shape[0], len(self._items))
            self.mapped_item_content = np.zeros((n, item_content.shape[1]), dtype=np.float64)
            self.mapped_item_content[:len(self._items)] = item_content[self._items.raw]
            self.item_content_dim = item_content.shape[-1]

        self.source_warm_user_idx, self.source_warm_item_idx = warm_user_idx, warm_item_idx      # :78-85
        self.source_cold_user_idx, self.source_cold_item_idx = cold_user_idx, cold_item_idx
        self.mapped_warm_user_idx = self.get_user_id_list(warm_user_idx)
        self.mapped_warm_item_idx = self.get_item_id_list(warm_item_idx)
        self.mapped_cold_user_idx = self.get_user_id_list(cold_user_idx)
        self.mapped_cold_item_idx = self.get_item_id_list(cold_item_idx)
        self._lazy = {}

    @classmethod
    def from_disk(cls, dataset: str, cold_object: str, root: str = ".") -> "ArrayDataBuilder":
        """The data part of ``Config.__init__`` (main.py:28-57) on the reference's own on-disk layout, unchanged:
        ``<root>/data/<dataset>/cold_<obj>/{warm_train,warm_val,warm_test,cold_<obj>_val,cold_<obj>_test,overall_val,
        overall_test}.csv`` + ``info_dict.pkl`` (data/convert.py:116-143) and ``<root>/data/<dataset>/<dataset>_<obj>_content.npy``."""
        import os
        import pickle
        if cold_object not in ("user", "item"):
            raise ValueError(f"cold_object must be 'user' or 'item', got {cold_object!r}")
        base = os.path.join(root, "data", dataset, f"cold_{cold_object}")
        L = lambda f: load_data_set(os.path.join(base, f))
        with open(os.path.join(base, "info_dict.pkl"), "rb") as f:
            info = pickle.load(f)
        content = np.load(os.path.join(root, "data", dataset, f"{dataset}_{cold_object}_content.npy"))
        uc, ic = (content, None) if cold_object == "user" else (None, content)
        return cls(L("warm_train.csv"), L("warm_val.csv"), L(f"cold_{cold_object}_val.csv"), L("overall_val.csv"), L("warm_test.csv"),
                   L(f"cold_{cold_object}_test.csv"), L("overall_test.csv"), info["user_num"], info["item_num"], info["warm_user"],
                   info["warm_item"], info["cold_user"], info["cold_item"], uc, ic)

    # ---- id tables (reference attribute names) -------------------------------------------------------
    @property
    def user(self) -> Dict[int, int]:
        return self._users.as_dict()

    @property
    def item(self) -> Dict[int, int]:
        return self._items.as_dict()

    @property
    def id2user(self) -> Dict[int, int]:
        return self._memo("id2user", lambda: dict(enumerate(self._users.raw.tolist())))

    @property
    def id2item(self) -> Dict[int, int]:
        return self._memo("id2item", lambda: dict(enumerate(self._items.raw.tolist())))

    def get_user_id(self, u):                                       # :277-281
        return int(self._users.lookup(np.asarray([u]), "user")[0])

    def get_item_id(self, i):                                       # :283-287
        return int(self._items.lookup(np.asarray([i]), "item")[0])

    def get_user_id_list(self, u_list):                             # :289-296
        return self._users.lookup(np.asarray(list(u_list) if not isinstance(u_list, np.ndarray) else u_list, dtype=np.int64), "user")

    def get_item_id_list(self, i_list):                             # :298-305
        return self._items.lookup(np.asarray(list(i_list) if not isinstance(i_list, np.ndarray) else i_list, dtype=np.int64), "item")

    def _memo(self, key, make):
        if key not in self._lazy:
            self._lazy[key] = make()
        return self._lazy[key]

    # ---- matrices (:220-275) --------------------------------------------------------------------------
    @property
    def ui_adj(self):
        def make():
            n = self.user_num + self.item_num
            ones = np.ones(len(self.train_user), dtype=np.float32)
            tmp = sp.csr_matrix((ones, (self.train_user, self.train_item + self.user_num)), shape=(n, n), dtype=np.float32)
            return tmp + tmp.T
        return self._memo("ui_adj", make)

    @staticmethod
    def normalize_graph_mat(adj_mat):                               # :236-254
        shape = adj_mat.get_shape()
        rowsum = np.array(adj_mat.sum(1)).flatten()
        d_inv = np.zeros_like(rowsum, dtype=np.float32)
        if shape[0] == shape[1]:
            np.power(rowsum, -0.5, out=d_inv, where=rowsum != 0)
            d_mat_inv = sp.diags(d_inv)
            return d_mat_inv.dot(adj_mat).dot(d_mat_inv)
        np.power(rowsum, -1, out=d_inv, where=rowsum != 0)
        return sp.diags(d_inv).dot(adj_mat)

    @property
    def norm_adj(self):
        return self._memo("norm_adj", lambda: self.normalize_graph_mat(self.ui_adj))

    @property
    def interaction_mat(self):                                      # :265-275
        return self._memo("interaction_mat", lambda: sp.csr_matrix(
            (np.ones(len(self.train_user), dtype=np.float32), (self.train_user, self.train_item)),
            shape=(self.user_num, self.item_num), dtype=np.float32))

    # ---- dict views of the reference (lazy; not used by the hot path) ----------------------------------
    def _dict_of_dict(self, split: str, by_item: bool = False):
        def make():
            out = defaultdict(dict)
            p = self._pairs[split]
            a, b = (p[:, 1], p[:, 0]) if by_item else (p[:, 0], p[:, 1])
            rows = getattr(self, "training_data" if split == "training" else f"{split}_data")
            ratings = [r[2] for r in rows] if (not isinstance(rows, np.ndarray) and len(rows) and len(rows[0]) > 2) else None
            for k, (x, y) in enumerate(zip(a.tolist(), b.tolist())):
                out[x][y] = ratings[k] if ratings is not None else 1.0
            return out
        return self._memo(("dd", split, by_item), make)

    training_set_u = property(lambda self: self._dict_of_dict("training"))
    training_set_i = property(lambda self: self._dict_of_dict("training", True))
    warm_valid_set = property(lambda self: self._dict_of_dict("warm_valid"))
    warm_test_set = property(lambda self: self._dict_of_dict("warm_test"))
    cold_valid_set = property(lambda self: self._dict_of_dict("cold_valid"))
    cold_test_set = property(lambda self: self._dict_of_dict("cold_test"))
    overall_valid_set = property(lambda self: self._dict_of_dict("overall_valid"))
    overall_test_set = property(lambda self: self._dict_of_dict("overall_test"))

    @property
    def training_set_uid(self):                                     # :108-114 (array of sets, by dense user id)
        def make():
            out = np.array([set() for _ in range(max(self.user_num, len(self._users)))])
            raw_items = self._pairs["training"][:, 1]
            order = np.argsort(self.train_user, kind="stable")
            bounds = np.searchsorted(self.train_user[order], np.arange(len(out) + 1))
            for uid in range(len(out)):
                if bounds[uid + 1] > bounds[uid]:
                    out[uid] = set(raw_items[order[bounds[uid]:bounds[uid + 1]]].tolist())
            return out
        return self._memo("training_set_uid", make)

    def training_size(self):                                        # :307-308: the whole id tables (cold entities included)
        return len(self._users), len(self._items), len(self.train_user)

    def _split_size(self, split):
        u, i = self._dense[split]
        return len(np.unique(u)), len(np.unique(i)), len(u)

    warm_valid_size = lambda self: self._split_size("warm_valid")
    warm_test_size = lambda self: self._split_size("warm_test")
    cold_valid_size = lambda self: self._split_size("cold_valid")
    cold_test_size = lambda self: self._split_size("cold_test")
    overall_valid_size = lambda self: self._split_size("overall_valid")
    overall_test_size = lambda self: self._split_size("overall_test")

    # ---- what the kernels consume --------------------------------------------------------------------
    def train_csr(self):
        """Train interactions as CSR over dense user ids: rowptr int64 [n_users+1], col int32 ascending, duplicates collapsed."""
        def make():
            n_u = max(self.user_num, len(self._users))
            n_i = max(self.item_num, len(self._items))
            keys = np.unique(self.train_user * n_i + self.train_item)
            rows = keys // n_i
            rowptr = np.zeros(n_u + 1, dtype=np.int64)
            np.cumsum(np.bincount(rows, minlength=n_u), out=rowptr[1:])
            return rowptr, (keys - rows * n_i).astype(np.int32)
        return self._memo("train_csr", make)

    def eval_arrays(self, split: str):
        """Eval users of a split in ground-truth dict order (first appearance, model/BaseRecommender.py:115), with the CSR of
        their train items (the mask, :117-128) and of their ground-truth items, dense ids ascending per row."""
        def make():
            u, i = self._dense[split]
            _, first = np.unique(u, return_index=True)
            uids = u[np.sort(first)]                                 # first-appearance order
            rank = np.full(max(self.user_num, len(self._users)), -1, dtype=np.int64)
            rank[uids] = np.arange(len(uids))
            n_i = max(self.item_num, len(self._items))
            gk = np.unique(rank[u] * n_i + i)                        # duplicates collapse like the dict does
            g_rows = gk // n_i
            gt_rowptr = np.zeros(len(uids) + 1, dtype=np.int64)
            np.cumsum(np.bincount(g_rows, minlength=len(uids)), out=gt_rowptr[1:])
            t_rowptr, t_col = self.train_csr()
            lens = t_rowptr[uids + 1] - t_rowptr[uids]
            mask_rowptr = np.zeros(len(uids) + 1, dtype=np.int64)
            np.cumsum(lens, out=mask_rowptr[1:])
            idx = np.repeat(t_rowptr[uids] - mask_rowptr[:-1], lens) + np.arange(mask_rowptr[-1])
            return dict(users=self._users.raw[uids], user_ids=uids.astype(np.int32), mask_rowptr=mask_rowptr,
                        mask_col=t_col[idx].astype(np.int32), gt_rowptr=gt_rowptr, gt_col=(gk - g_rows * n_i).astype(np.int32))
        return self._memo(("eval", split), make)

    def split_of(self, data_set) -> Optional[str]:
        """Name of the eval split whose (memoised) dict view ``data_set`` IS — how ``_evaluate(data_set, data_type)``, which
        only gets the dict, finds its way back to the arrays.  None for any other object."""
        for split in SPLITS:
            if split != "training" and self._lazy.get(("dd", split, False)) is data_set:
                return split
        return None

    def eval_plan(self, split: str, data_type: str, cold_object: str, device):
        """The scorer's ``EvalPlan`` for a split ('warm_test', 'overall_valid', ...) without any per-user Python work."""
        import torch
        from .scoring import EvalPlan, flag_exclude_for
        a = self.eval_arrays(split)
        t = lambda x: torch.from_numpy(x).to(device)
        return EvalPlan(a["users"].tolist(), t(a["user_ids"]), t(a["mask_rowptr"]), t(a["mask_col"]), t(a["gt_rowptr"]), t(a["gt_col"]),
                        flag_exclude_for(cold_object, data_type))

    def item_flags(self) -> np.ndarray:
        """bit0 = cold item, bit1 = warm item (model/BaseRecommender.py:130-143's column masks as one byte per item)."""
        flags = np.zeros(max(self.item_num, len(self._items)), dtype=np.uint8)
        flags[np.asarray(self.mapped_cold_item_idx, dtype=np.int64)] |= 1
        flags[np.asarray(self.mapped_warm_item_idx, dtype=np.int64)] |= 2
        return flags

    def graph(self, device):
        """Normalised bipartite adjacency on the device (``CsrGraph``), built by ``bipartite_norm_csr`` from the id arrays."""
        import torch
        from .graph import bipartite_norm_csr
        return bipartite_norm_csr(torch.from_numpy(self.train_user).to(device), torch.from_numpy(self.train_item).to(device),
                                  self.user_num, self.item_num)

    def sampler(self, device, seed: int = 2024):
        import torch
        from .training import PairwiseSampler
        return PairwiseSampler(torch.from_numpy(self.train_user).to(device), torch.from_numpy(self.train_item).to(device),
                               max(self.user_num, len(self._users)), len(self._items), seed)
