"""CPU: the array-based data builder against the reference's ColdStartDataBuilder — through the golden
id tables the reference produced (tests/golden/eval_*.npz) and through the oracle's line-by-line restatement of it."""
import numpy as np
import pytest

from coldrec_b200.databuilder import ArrayDataBuilder
from oracle import coldrec_oracle as O
from tests.helpers import builder_args, load_golden


@pytest.mark.parametrize("name", ["eval_item", "eval_user", "eval_tiny"])
def test_id_tables_match_the_reference(name):
    g = load_golden(name)
    data = ArrayDataBuilder(*builder_args(g))
    assert [data.id2user[i] for i in range(len(data.user))] == g["id2user"].tolist()
    assert [data.id2item[i] for i in range(len(data.item))] == g["id2item"].tolist()
    for k in ("mapped_cold_item_idx", "mapped_warm_item_idx", "mapped_cold_user_idx", "mapped_warm_user_idx"):
        assert np.array_equal(np.asarray(getattr(data, k)), g[k]), k
    gg = load_golden("graph") if name == "eval_item" else None
    if gg is not None:      # the reference's own adjacency
        adj = data.norm_adj.tocsr(); adj.sort_indices()
        assert np.array_equal(adj.indptr, gg["adj_indptr"]) and np.array_equal(adj.indices, gg["adj_indices"])
        assert np.allclose(adj.data, gg["adj_data"], rtol=1e-6, atol=0)
        assert np.array_equal(data.train_user, gg["train_u"]) and np.array_equal(data.train_item, gg["train_i"])


def _random_args(seed, n_users=60, n_items=45, n_inter=900, with_content=True):
    rng = np.random.default_rng(seed)
    raw_u = rng.permutation(500)[:n_users] + 7          # sparse, unordered raw ids
    raw_i = rng.permutation(400)[:n_items] + 3
    pairs = np.stack([raw_u[rng.integers(0, n_users, n_inter)], raw_i[rng.integers(0, n_items, n_inter)]], 1)   # with duplicates
    cut = np.cumsum([0.6, 0.07, 0.07, 0.06, 0.06, 0.07])
    parts = np.split(pairs, (cut * n_inter).astype(int))
    rows = [[[int(u), int(i), float(1 + (k % 3))] for k, (u, i) in enumerate(p)] for p in parts]
    train, wv, wt, cv, ct, ov, ot = rows
    seen_u = sorted({r[0] for p in rows for r in p}); seen_i = sorted({r[1] for p in rows for r in p})
    half_u, half_i = len(seen_u) // 2, len(seen_i) // 2
    ucont = rng.standard_normal((520, 5)) if with_content else None
    icont = rng.standard_normal((420, 4)) if with_content else None
    return (train, wv, cv, ov, wt, ct, ot, len(seen_u) + 3, len(seen_i) + 2, seen_u[:half_u], seen_i[:half_i], seen_u[half_u:],
            seen_i[half_i:], ucont, icont)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_equals_the_restated_reference_builder(seed):
    args = _random_args(seed)
    ref, got = O.OracleData(*args), ArrayDataBuilder(*args)
    assert got.user == ref.user and got.item == ref.item
    assert got.id2user == ref.id2user and got.id2item == ref.id2item
    for k in ("mapped_cold_item_idx", "mapped_warm_item_idx", "mapped_cold_user_idx", "mapped_warm_user_idx"):
        assert np.array_equal(getattr(got, k), getattr(ref, k))
    for k in ("training_set_u", "training_set_i", "warm_valid_set", "warm_test_set", "cold_valid_set", "cold_test_set",
              "overall_valid_set", "overall_test_set"):
        a, b = getattr(got, k), getattr(ref, k)
        assert {u: dict(v) for u, v in a.items()} == {u: dict(v) for u, v in b.items()}, k
        assert list(a.keys()) == list(b.keys()), k             # eval-user order is the dict order (BaseRecommender.py:115)
    assert (got.ui_adj != ref.ui_adj).nnz == 0
    d = (got.norm_adj - ref.norm_adj)
    assert abs(d).max() <= 1e-7
    n_u, n_i = len(ref.user), len(ref.item)
    assert np.array_equal(got.mapped_user_content[:n_u], ref.mapped_user_content[:n_u])
    assert np.array_equal(got.mapped_item_content[:n_i], ref.mapped_item_content[:n_i])
    assert got.get_user_id(ref.id2user[3]) == 3 and got.get_item_id(ref.id2item[5]) == 5


@pytest.mark.parametrize("split", ["warm_test", "cold_test", "overall_valid"])
def test_eval_arrays_equal_the_per_user_construction(split):
    args = _random_args(5)
    ref, got = O.OracleData(*args), ArrayDataBuilder(*args)
    a = got.eval_arrays(split)
    gt = getattr(ref, f"{split}_set")
    users = list(gt.keys())
    assert a["users"].tolist() == users
    assert a["user_ids"].tolist() == [ref.user[u] for u in users]
    for j, u in enumerate(users):
        want_mask = sorted(ref.item[i] for i in ref.training_set_u[u]) if u in ref.training_set_u else []
        want_gt = sorted(ref.item[i] for i in gt[u])
        assert a["mask_col"][a["mask_rowptr"][j]:a["mask_rowptr"][j + 1]].tolist() == want_mask
        assert a["gt_col"][a["gt_rowptr"][j]:a["gt_rowptr"][j + 1]].tolist() == want_gt


def test_train_csr_flags_sets_and_errors():
    args = _random_args(9, with_content=False)
    ref, got = O.OracleData(*args), ArrayDataBuilder(*args)
    rowptr, col = got.train_csr()
    m = got.interaction_mat.tocsr(); m.sum_duplicates(); m.sort_indices()
    assert np.array_equal(rowptr[:m.shape[0] + 1], m.indptr) and np.array_equal(col, m.indices)
    flags = got.item_flags()
    assert set(np.nonzero(flags & 1)[0]) == set(ref.mapped_cold_item_idx.tolist())
    assert set(np.nonzero(flags & 2)[0]) == set(ref.mapped_warm_item_idx.tolist())
    uid_sets = got.training_set_uid
    for u, uid in ref.user.items():
        assert uid_sets[uid] == set(ref.training_set_u[u]) if u in ref.training_set_u else uid_sets[uid] == set()
    assert got.training_size() == (len(ref.user), len(ref.item), len(args[0]))      # util/databuilder.py:307-308
    with pytest.raises(Exception, match="user 99999 not in current id table"):
        got.get_user_id(99999)
    with pytest.raises(Exception, match="item 12345 not in current id table"):
        got.get_item_id_list([args[0][0][1], 12345])


def test_accepts_arrays_and_empty_splits():
    args = list(_random_args(3, with_content=False))
    as_arrays = [np.asarray([(r[0], r[1]) for r in a], dtype=np.int64).reshape(-1, 2) for a in args[:7]]
    a, b = ArrayDataBuilder(*args), ArrayDataBuilder(*as_arrays, *args[7:])
    assert a.user == b.user and a.item == b.item and np.array_equal(a.train_csr()[1], b.train_csr()[1])
    args[2] = []                     # no cold-valid interactions at all
    c = ArrayDataBuilder(*args)
    e = c.eval_arrays("cold_valid")
    assert len(e["user_ids"]) == 0 and e["mask_rowptr"].tolist() == [0] and e["gt_rowptr"].tolist() == [0]


def test_from_disk_reads_the_reference_layout_unchanged():
    """tests/golden/disk/ was written by the reference's own data/split.py + data/convert.py; expect.npz is what its
    DataLoader + ColdStartDataBuilder make of it (oracle/make_golden.py --only disk)."""
    import os
    import scipy.sparse as sp
    from coldrec_b200.databuilder import ArrayDataBuilder, load_data_set
    from tests.helpers import GOLDEN
    root = os.path.join(GOLDEN, "disk")
    e = dict(np.load(os.path.join(root, "expect.npz")))
    pairs = load_data_set(os.path.join(root, "data", "syntiny", "cold_item", "warm_train.csv"))
    assert pairs.dtype == np.int64 and pairs.shape == (int(e["n_train"]), 2)
    b = ArrayDataBuilder.from_disk("syntiny", "item", root=root)
    assert (b.user_num, b.item_num) == (int(e["user_num"]), int(e["item_num"]))
    assert np.array_equal(np.array([b.id2user[i] for i in range(len(b.user))]), e["id2user"])
    assert np.array_equal(np.array([b.id2item[i] for i in range(len(b.item))]), e["id2item"])
    for k in ("mapped_warm_item_idx", "mapped_cold_item_idx", "mapped_warm_user_idx", "mapped_cold_user_idx"):
        assert np.array_equal(np.asarray(getattr(b, k)), e[k]), k
    assert np.array_equal(b.mapped_item_content[:len(b.item)], e["mapped_item_content"])
    adj = b.norm_adj.tocsr(); adj.sort_indices()
    ref = sp.csr_matrix((e["norm_adj_data"], e["norm_adj_indices"], e["norm_adj_indptr"]), shape=adj.shape); ref.sort_indices()
    assert np.array_equal(adj.indptr, ref.indptr) and np.array_equal(adj.indices, ref.indices)
    assert np.abs(adj.data - ref.data).max() <= 1e-7
    with pytest.raises(ValueError):
        ArrayDataBuilder.from_disk("syntiny", "both", root=root)
    with pytest.raises(FileNotFoundError):
        ArrayDataBuilder.from_disk("nosuch", "item", root=root)


def test_mask_rows_remap_for_compacted_tables():
    """FullRankScorer._mask_in_rows_of: the train-mask CSR re-expressed in row numbers of a compacted item table — entries whose
    item was compacted away are dropped, the others become their row number, per-row order kept, memoised per (plan, selection)."""
    import torch
    from coldrec_b200.scoring import EvalPlan, FullRankScorer
    rng = np.random.default_rng(3)
    n_items, n_q = 500, 40
    rows = [np.sort(rng.choice(n_items, int(rng.integers(0, 30)), replace=False)) for _ in range(n_q)]
    rows[5] = np.zeros(0, dtype=np.int64)
    rowptr = np.zeros(n_q + 1, dtype=np.int64); np.cumsum([len(r) for r in rows], out=rowptr[1:])
    col = np.concatenate(rows).astype(np.int32)
    plan = EvalPlan(None, torch.arange(n_q, dtype=torch.int32), torch.from_numpy(rowptr), torch.from_numpy(col), torch.zeros(n_q + 1, dtype=torch.int64),
                    torch.zeros(0, dtype=torch.int32), 1)
    keep = rng.random(n_items) < 0.6
    gids = torch.from_numpy(np.nonzero(keep)[0].astype(np.int32))
    sc = FullRankScorer.__new__(FullRankScorer)
    sc._remap_cache = {}
    rp, c = sc._mask_in_rows_of(plan, gids, "k")
    pos_of = {int(g): j for j, g in enumerate(gids.tolist())}
    for j in range(n_q):
        want = [pos_of[int(g)] for g in rows[j] if int(g) in pos_of]
        assert c[int(rp[j]):int(rp[j + 1])].tolist() == want
    assert sc._mask_in_rows_of(plan, gids, "k")[1] is c            # memoised
    assert sc._mask_in_rows_of(plan, gids, "other")[1] is not c
    gids2 = gids[:-3].clone()                                       # a rebuilt id list (new flags) under the same key: never the old remap
    rp2, c2 = sc._mask_in_rows_of(plan, gids2, "k")
    assert c2 is not c and int(c2.max()) < gids2.numel()
