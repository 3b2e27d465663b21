"""GPU parity: the CUDA path (through the C ABI) vs. the oracle / the reference's golden vectors.

Tolerances (BASELINE.json north_star):
  * top-K ids identical except at score ties within 1e-5; scores within 1e-5 (relative to max(1,|s|))
  * Recall/NDCG within 1e-6 of the unrounded formulas and the printed (rounded) strings equal
  * propagated / generated embeddings within 1e-5 norm-wise: max|d| <= 1e-5 * max|ref|
"""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

from oracle import coldrec_oracle as O
from tests.helpers import builder_args, load_golden, rec_from_golden, t

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def cu(x, dtype=None):
    x = torch.from_numpy(np.ascontiguousarray(x)) if isinstance(x, np.ndarray) else x
    if dtype is not None:
        x = x.to(dtype)
    return x.to(DEV)


def assert_normwise(got: torch.Tensor, ref, tol=1e-5):
    ref = ref if isinstance(ref, np.ndarray) else ref.numpy()
    got = got.cpu().numpy()
    assert got.shape == ref.shape
    err = np.abs(got - ref).max() if ref.size else 0.0
    assert err <= tol * max(np.abs(ref).max(), 1e-30), f"max abs err {err} vs scale {np.abs(ref).max()}"


# ---------------------------------------------------------------------------------------------- SpMM
def _rand_csr(rng, n_rows, n_cols, density_rows, long_rows=()):
    rows, cols = [], []
    for r in range(n_rows):
        k = long_rows[r] if r in long_rows else int(rng.integers(0, density_rows))
        c = rng.choice(n_cols, size=min(k, n_cols), replace=False)
        rows += [r] * len(c)
        cols += c.tolist()
    vals = rng.standard_normal(len(rows)).astype(np.float32)
    return sp.csr_matrix((vals, (rows, cols)), shape=(n_rows, n_cols))


@pytest.mark.parametrize("d", [32, 64, 128, 256, 48, 96, 200])
def test_spmm_matches_torch_sparse_mm(d):
    from coldrec_b200 import CsrGraph
    rng = np.random.default_rng(d)
    A = _rand_csr(rng, 300, 257, 40, long_rows={5: 257, 17: 0, 299: 200})
    X = torch.from_numpy(rng.standard_normal((257, d)).astype(np.float32))
    ref = torch.sparse.mm(O.sparse_mat_to_torch(A), X)
    g = CsrGraph.from_scipy(A, DEV)
    y = g.spmm(cu(X), Y=torch.empty(300, d, device=DEV))
    assert_normwise(y, ref)


def test_spmm_long_rows_split_path_and_epilogue():
    from coldrec_b200 import CsrGraph
    rng = np.random.default_rng(7)
    n = 6000
    A = _rand_csr(rng, 64, n, 30, long_rows={0: 5000, 3: 513, 7: 512, 9: 1537, 63: 2048})
    X = torch.from_numpy(rng.standard_normal((n, 64)).astype(np.float32))
    acc0 = torch.from_numpy(rng.standard_normal((64, 64)).astype(np.float32))
    ref_y = torch.sparse.mm(O.sparse_mat_to_torch(A), X)
    g = CsrGraph.from_scipy(A, DEV)
    y, acc = torch.empty(64, 64, device=DEV), cu(acc0.clone())
    g.spmm(cu(X), Y=y, acc=acc, acc_beta=1.0, acc_div=3.0)
    assert_normwise(y, ref_y)
    assert_normwise(acc, (acc0 + ref_y) / 3.0)
    # all-ones values (val == NULL) and the no-plan path give the same numbers
    from coldrec_b200 import ops
    A1 = A.copy(); A1.data[:] = 1.0
    ref1 = torch.sparse.mm(O.sparse_mat_to_torch(A1), X)
    y1 = ops.spmm(g.rowptr, g.col, None, cu(X), Y=torch.empty(64, 64, device=DEV), plan=None)
    assert_normwise(y1, ref1)
    y2 = ops.spmm(g.rowptr, g.col, None, cu(X), Y=torch.empty(64, 64, device=DEV), plan=g.plan(64))
    assert_normwise(y2, ref1)
    # split path is deterministic
    y3 = ops.spmm(g.rowptr, g.col, None, cu(X), Y=torch.empty(64, 64, device=DEV), plan=g.plan(64))
    assert torch.equal(y2, y3)


def test_spmm_empty_and_linearity():
    from coldrec_b200 import CsrGraph
    rng = np.random.default_rng(3)
    A = _rand_csr(rng, 500, 400, 12)
    g = CsrGraph.from_scipy(A, DEV)
    X1, X2 = torch.randn(400, 64, device=DEV), torch.randn(400, 64, device=DEV)
    y1 = g.spmm(X1, Y=torch.empty(500, 64, device=DEV))
    y2 = g.spmm(X2, Y=torch.empty(500, 64, device=DEV))
    y12 = g.spmm(X1 + 2 * X2, Y=torch.empty(500, 64, device=DEV))
    assert torch.allclose(y12, y1 + 2 * y2, atol=1e-4)
    E = sp.csr_matrix((0, 400), dtype=np.float32)
    ge = CsrGraph.from_scipy(E, DEV)
    assert ge.spmm(X1, Y=torch.empty(0, 64, device=DEV)).shape == (0, 64)


def _golden_graph():
    g = load_golden("graph")
    n = int(g["user_num"]) + int(g["item_num"])
    return g, sp.csr_matrix((g["adj_data"], g["adj_indices"], g["adj_indptr"]), shape=(n, n))


@pytest.mark.parametrize("L", [1, 2, 3])
def test_lightgcn_propagate_vs_reference_golden(L):
    from coldrec_b200 import CsrGraph, propagate
    g, A = _golden_graph()
    G = CsrGraph.from_scipy(A, DEV)
    u, i = propagate(G, cu(g["E0_user"]), cu(g["E0_item"]), L)
    assert_normwise(u, g[f"lgcn_L{L}_user"])
    assert_normwise(i, g[f"lgcn_L{L}_item"])
    # allocation-free form (a trainer evaluates every epoch): same bits, results are views of the caller's buffers
    from coldrec_b200 import PropagationBuffers
    bufs = PropagationBuffers(A.shape[0], 64, DEV, with_ego=True)
    for _ in range(2):
        u2, i2 = propagate(G, cu(g["E0_user"]), cu(g["E0_item"]), L, buffers=bufs)
        assert torch.equal(u2, u) and torch.equal(i2, i) and u2.data_ptr() == bufs.acc.data_ptr()


def test_simgcl_ngcf_and_layer_list_vs_reference_golden():
    from coldrec_b200 import CsrGraph, propagate, propagate_ngcf
    g, A = _golden_graph()
    G = CsrGraph.from_scipy(A, DEV)
    u, i = propagate(G, cu(g["E0_user"]), cu(g["E0_item"]), 3, include_ego=False)
    assert_normwise(u, g["simgcl_L3_user"]); assert_normwise(i, g["simgcl_L3_item"])
    u, i, layers = propagate(G, cu(g["E0_user"]), cu(g["E0_item"]), 3, return_layers=True)
    assert_normwise(u, g["lgcn_L3_user"]) and len(layers) == 4
    _, _, ref_layers = O.propagate(A, t(g["E0_user"]), t(g["E0_item"]), 3, return_layers=True)
    for a, b in zip(layers, ref_layers):
        assert_normwise(a, b)
    Wgc = [(cu(g[f"ngcf_Wgc{l}_w"]), cu(g[f"ngcf_Wgc{l}_b"])) for l in range(2)]
    Wbi = [(cu(g[f"ngcf_Wbi{l}_w"]), cu(g[f"ngcf_Wbi{l}_b"])) for l in range(2)]
    u, i = propagate_ngcf(G, cu(g["E0_user"]), cu(g["E0_item"]), Wgc, Wbi)
    assert_normwise(u, g["ngcf_L2_user"]); assert_normwise(i, g["ngcf_L2_item"])


def test_device_adjacency_builder_vs_reference_golden():
    from coldrec_b200 import bipartite_norm_csr
    g, A = _golden_graph()
    G = bipartite_norm_csr(cu(g["train_u"]), cu(g["train_i"]), int(g["user_num"]), int(g["item_num"]))
    assert np.array_equal(G.rowptr.cpu().numpy(), g["adj_indptr"])
    assert np.array_equal(G.col.cpu().numpy(), g["adj_indices"])
    assert np.allclose(G.val.cpu().numpy(), g["adj_data"], rtol=1e-6, atol=0)   # pow(-0.5) differs by an ulp between numpy and CUDA


# ---------------------------------------------------------------------------------------------- scoring
class _Args:
    def __init__(self, cold_object, bs=4096):
        self.topN, self.model, self.dataset, self.emb_size, self.epochs, self.bs = "10,20", "MF", "syn", 64, 0, bs
        self.lr, self.reg, self.early_stop, self.eval_every, self.cold_object = 1e-3, 1e-4, 0, 1, cold_object


class _Cfg:
    def __init__(self, data, cold_object):
        self.args, self.data, self.device = _Args(cold_object), data, torch.device(DEV)


def _trainer(data, cold_object, precision, base=None):
    from coldrec_b200 import BaseColdStartTrainer
    bases = (base, BaseColdStartTrainer) if base else (BaseColdStartTrainer,)

    class T(*bases):
        def train(self): pass
        def save(self): pass
    tr = T(_Cfg(data, cold_object))
    tr.score_precision = precision
    return tr


def _exact_scores_fn(data, score_rows, users, typ, cold_object):
    """Oracle (masked) scores of given dense item ids for eval row j."""
    cm = None
    if cold_object == "item" and typ in ("warm", "cold"):
        cm = set(np.asarray(data.mapped_cold_item_idx if typ == "warm" else data.mapped_warm_item_idx).tolist())

    def fn(j, ids):
        u = users[j]
        tr = {data.item[i] for i in data.training_set_u[u]} if u in data.training_set_u else set()
        row = score_rows(j)
        return [O.MASK_SENTINEL if (int(i) in tr or (cm is not None and int(i) in cm)) else float(row[int(i)]) for i in ids]
    return fn


PRECISIONS = [0, 1]


@pytest.mark.parametrize("precision", PRECISIONS, ids=["exact", "tf32"])
@pytest.mark.parametrize("name", ["eval_item", "eval_user", "eval_tiny"])
@pytest.mark.parametrize("typ", ["all", "cold", "warm"])
def test_evaluate_and_metrics_vs_reference_golden(name, typ, precision):
    g = load_golden(name)
    data = O.OracleData(*builder_args(g))
    cold_object = str(g["cold_object"])
    tr = _trainer(data, cold_object, precision)
    tr.user_emb, tr.item_emb = cu(g["user_emb"]), cu(g["item_emb"])
    rec = tr.test(typ)
    p = f"mf_test_{typ}"
    users = g[f"{p}_users"].tolist()
    assert rec.plan.users == users
    ue, ie = t(g["user_emb"]), t(g["item_emb"])
    uid = data.get_user_id_list(users)
    score_rows = lambda j: (ue[uid[j]] @ ie.T).numpy()
    O.check_topk_parity(g[f"{p}_scores"], g[f"{p}_dense_ids"], rec.scores.cpu().numpy(), rec.ids.cpu().numpy().astype(np.int64),
                        _exact_scores_fn(data, score_rows, users, typ, cold_object))
    # dict form: raw ids, python floats, sorted descending
    first = rec[users[0]]
    assert len(first) == 20 and all(a[1] >= b[1] for a, b in zip(first, first[1:]))
    # metrics: golden strings (rounded) and the unrounded oracle within 1e-6.  Lists that contain masked
    # ids (fewer than K unmasked candidates) may legitimately differ in *which* masked ids they show.
    gt = {"all": data.overall_test_set, "warm": data.warm_test_set, "cold": data.cold_test_set}[typ]
    from coldrec_b200.evaluator import device_metrics
    if (g[f"{p}_scores"] > -1e8).all():
        measure, perf = tr._ranking_evaluation(gt, rec, [10, 20])
        assert measure == g[f"{p}_measure"].tolist()
        unr = device_metrics(rec.ids, rec.plan.gt_rowptr, rec.plan.gt_col, [10, 20], rounded=False)
        want = O.ranking_metrics_unrounded(gt, rec_from_golden(g, p), [10, 20])
        assert np.allclose(unr, want, atol=1e-6, rtol=0)
    # a plain reference-style dict goes through the same device reduction
    measure2, _ = tr._ranking_evaluation(gt, rec_from_golden(g, p), [10, 20])
    assert measure2 == g[f"{p}_measure"].tolist()


@pytest.mark.parametrize("precision", PRECISIONS, ids=["exact", "tf32"])
@pytest.mark.parametrize("typ", ["all", "cold"])
def test_aldi_dual_tables_vs_reference_golden(typ, precision):
    from coldrec_b200 import AldiScoreTables
    g = load_golden("eval_item")
    data = O.OracleData(*builder_args(g))
    tr = _trainer(data, "item", precision, base=AldiScoreTables)
    tr.warm_user_emb, tr.cold_user_emb, tr.item_emb = cu(g["user_emb"]), cu(g["aldi_cold_user_emb"]), cu(g["item_emb"])
    rec = tr.test(typ)
    p = f"aldi_test_{typ}"
    users = g[f"{p}_users"].tolist()
    uid = data.get_user_id_list(users)
    fn = O.score_aldi(t(g["user_emb"]), t(g["aldi_cold_user_emb"]), t(g["item_emb"]), data.mapped_warm_item_idx, data.mapped_cold_item_idx)
    score_rows = lambda j: fn(torch.tensor([uid[j]]))[0].numpy()
    O.check_topk_parity(g[f"{p}_scores"], g[f"{p}_dense_ids"], rec.scores.cpu().numpy(), rec.ids.cpu().numpy().astype(np.int64),
                        _exact_scores_fn(data, score_rows, users, typ, "item"))


def test_reevaluation_after_item_table_is_freed_and_reallocated():
    """trainers rebind ``item_emb`` to a fresh tensor every epoch and the caching allocator hands the
    same address (version 0) back two epochs later; a compaction cache keyed on (address, version) then returned an old
    epoch's rows.  Evaluate ALDI-style (three compacted item groups) over six epochs of freed-and-reallocated tables with
    different contents and compare every epoch with the oracle."""
    from coldrec_b200 import AldiScoreTables
    g = load_golden("eval_item")
    data = O.OracleData(*builder_args(g))
    tr = _trainer(data, "item", PRECISIONS[1], base=AldiScoreTables)
    rng = np.random.default_rng(5)
    tr.item_emb = None
    for epoch in range(6):
        I = (rng.standard_normal(g["item_emb"].shape) * 0.1).astype(np.float32)
        Uw = (rng.standard_normal(g["user_emb"].shape) * 0.1).astype(np.float32)
        Uc = (rng.standard_normal(g["user_emb"].shape) * 0.1).astype(np.float32)
        if epoch % 2 == 1:                                   # same storage, same _version, new contents (ALDI's `.data[...] = ...`):
            key = (tr.item_emb.data_ptr(), tr.item_emb._version)
            tr.item_emb.data.copy_(cu(I))                    # exactly the state an (address, version) key cannot tell apart
            assert key == (tr.item_emb.data_ptr(), tr.item_emb._version)
        else:                                                # a rebinding trainer: last epoch's block freed, a new one allocated
            tr.item_emb = None
            tr.item_emb = cu(I).clone()
        tr.warm_user_emb, tr.cold_user_emb = cu(Uw), cu(Uc)
        for typ in ("cold", "warm"):
            rec = tr.test(typ)
            users = list(getattr(data, f"{typ}_test_set").keys())
            uid = data.get_user_id_list(users)
            fn = O.score_aldi(t(Uw), t(Uc), t(I), data.mapped_warm_item_idx, data.mapped_cold_item_idx)
            score_rows = lambda j: fn(torch.tensor([uid[j]]))[0].numpy()
            exact = _exact_scores_fn(data, score_rows, users, typ, "item")
            got_s, got_i = rec.scores.cpu().numpy(), rec.ids.cpu().numpy().astype(np.int64)
            for j in range(len(users)):
                want = np.asarray(exact(j, got_i[j]), dtype=np.float32)
                assert np.allclose(got_s[j], want, atol=1e-5), f"epoch {epoch} {typ} user {j}: scores are not this epoch's"


@pytest.mark.parametrize("precision", PRECISIONS, ids=["exact", "tf32"])
def test_vbpr_two_products_vs_reference_golden(precision):
    from coldrec_b200 import TwoProductScoreTables
    g = load_golden("eval_item")
    data = O.OracleData(*builder_args(g))
    tr = _trainer(data, "item", precision, base=TwoProductScoreTables)
    tr.user_emb_main, tr.item_emb_main = cu(g["user_emb"]), cu(g["item_emb"])
    tr.user_emb_aux, tr.item_emb_aux = cu(g["vbpr_user_aux"]), cu(g["vbpr_item_aux"])
    rec = tr.test("all")
    p = "vbpr_test_all"
    users = g[f"{p}_users"].tolist()
    uid = data.get_user_id_list(users)
    fn = O.score_vbpr(t(g["user_emb"]), t(g["item_emb"]), t(g["vbpr_user_aux"]), t(g["vbpr_item_aux"]))
    score_rows = lambda j: fn(torch.tensor([uid[j]]))[0].numpy()
    O.check_topk_parity(g[f"{p}_scores"], g[f"{p}_dense_ids"], rec.scores.cpu().numpy(), rec.ids.cpu().numpy().astype(np.int64),
                        _exact_scores_fn(data, score_rows, users, "all", "item"))


def _synthetic_scoring_case(seed, n_users, n_items, n_q, d, mask_per_user, dup_frac=0.0):
    rng = np.random.default_rng(seed)
    U = (rng.standard_normal((n_users, d)) * 0.125).astype(np.float32)
    I = (rng.standard_normal((n_items, d)) * 0.125).astype(np.float32)
    if dup_frac:
        nd = int(n_items * dup_frac)
        I[rng.choice(n_items, nd, replace=False)] = I[rng.choice(n_items, nd, replace=False)]   # exact score ties
    uids = rng.choice(n_users, n_q, replace=False).astype(np.int32)
    rows = [np.sort(rng.choice(n_items, size=int(rng.integers(0, mask_per_user + 1)), replace=False)) for _ in range(n_q)]
    rowptr = np.zeros(n_q + 1, dtype=np.int64); np.cumsum([len(r) for r in rows], out=rowptr[1:])
    col = np.concatenate(rows).astype(np.int32) if rowptr[-1] else np.zeros(0, np.int32)
    flags = (rng.random(n_items) < 0.2).astype(np.uint8) * 1
    flags[flags == 0] = 2
    return U, I, uids, rowptr, col, flags


@pytest.mark.parametrize("precision", PRECISIONS, ids=["exact", "tf32"])
@pytest.mark.parametrize("shape", [(500, 3000, 333, 64, 40, 0.0), (900, 20000, 700, 64, 120, 0.02), (300, 5000, 300, 128, 30, 0.0),
                                   (64, 40, 64, 64, 10, 0.0), (200, 70000, 37, 64, 200, 0.01),
                                   # widths zero-padded onto the 64 / 128 tensor-core instantiations, and one beyond them (FFMA kernel)
                                   (400, 9000, 300, 32, 30, 0.0), (300, 6000, 256, 96, 20, 0.01), (200, 4000, 150, 48, 10, 0.0),
                                   (150, 3000, 100, 200, 10, 0.0)])
def test_score_topk_vs_oracle_synthetic(shape, precision):
    from coldrec_b200 import ops
    n_users, n_items, n_q, d, mpu, dup = shape
    U, I, uids, rowptr, col, flags = _synthetic_scoring_case(sum(shape[:4]), n_users, n_items, n_q, d, mpu, dup)
    for excl in (0, 1, 2):
        col_mask = None if excl == 0 else np.nonzero(flags & excl)[0]
        ref_s, ref_i = O.evaluate_topk_dense(O.score_mf(t(U), t(I)), uids, rowptr, col.astype(np.int64), col_mask, 20, 256)
        s, i, nref = ops.score_topk(cu(U), cu(I), 20, user_ids=cu(uids), mask_rowptr=cu(rowptr), mask_col=cu(col),
                                    item_flags=cu(flags) if excl else None, flag_exclude=excl, precision=precision)
        Ut, It = t(U), t(I)
        masked_sets = [set(col[rowptr[j]:rowptr[j + 1]].tolist()) for j in range(n_q)]
        cm = set() if col_mask is None else set(col_mask.tolist())

        def exact(j, ids):
            row = (Ut[uids[j]] @ It.T).numpy()
            return [O.MASK_SENTINEL if (int(x) in masked_sets[j] or int(x) in cm) else float(row[int(x)]) for x in ids]
        O.check_topk_parity(ref_s, ref_i, s.cpu().numpy(), i.cpu().numpy().astype(np.int64), exact)
        si = s.cpu().numpy()
        assert (np.diff(si, axis=1) <= 0).all(), "scores must be sorted descending"


@pytest.mark.parametrize("d", [64, 128])
@pytest.mark.parametrize("seed_tiles", [4, 16, 128])
@pytest.mark.parametrize("K", [20, 50])
def test_score_topk_seed_phase_vs_oracle(seed_tiles, K, d, monkeypatch):
    """The threshold seed phase of the tcgen05 scorer (first T0 tiles swept twice).  It switches on when a unit sweeps at
    least 8*T0 tiles; CR_TC_SEED_TILES (read per call) lowers T0 so that oracle-sized cases exercise it: masked and flagged
    items inside the seed tiles, 2 % duplicated item rows (exact ties at the seed threshold), K = 50 (KSEL = 64).  d = 128 is
    the VBPR / AMR instantiation of the same kernel (64-item tiles, four TMA boxes per tile, 2 x 16 MMAs)."""
    from coldrec_b200 import ops
    monkeypatch.setenv("CR_TC_SEED_TILES", str(seed_tiles))
    n_users, n_items, n_q = 2000, 110000 if seed_tiles == 128 else 30000, 1200
    U, I, uids, rowptr, col, flags = _synthetic_scoring_case(4242 + seed_tiles + K, n_users, n_items, n_q, d, 300, 0.02)
    col_head = np.sort(np.random.default_rng(K).choice(400, 60, replace=False)).astype(np.int32)   # masks concentrated in the seed tiles
    rows = [np.union1d(col[rowptr[j]:rowptr[j + 1]], col_head).astype(np.int32) for j in range(n_q)]
    rowptr = np.zeros(n_q + 1, dtype=np.int64); np.cumsum([len(r) for r in rows], out=rowptr[1:])
    col = np.concatenate(rows).astype(np.int32)
    for excl in (0, 2):
        col_mask = None if excl == 0 else np.nonzero(flags & excl)[0]
        ref_s, ref_i = O.evaluate_topk_dense(O.score_mf(t(U), t(I)), uids, rowptr, col.astype(np.int64), col_mask, K, 256)
        s, i, nref = ops.score_topk(cu(U), cu(I), K, user_ids=cu(uids), mask_rowptr=cu(rowptr), mask_col=cu(col),
                                    item_flags=cu(flags) if excl else None, flag_exclude=excl, precision=ops.SCORE_TF32_CHECKED)
        Ut, It = t(U), t(I)
        cm = set() if col_mask is None else set(col_mask.tolist())

        def exact(j, ids):
            masked = set(col[rowptr[j]:rowptr[j + 1]].tolist())
            row = (Ut[uids[j]] @ It.T).numpy()
            return [O.MASK_SENTINEL if (int(x) in masked or int(x) in cm) else float(row[int(x)]) for x in ids]
        O.check_topk_parity(ref_s, ref_i, s.cpu().numpy(), i.cpu().numpy().astype(np.int64), exact)
        assert int(nref.item()) < n_q // 20, "the TF32 margin proof should hold for almost every query"


def test_score_topk_degenerate_tables_all_scores_equal():
    """All-zero tables: every score ties at 0, the seed threshold sits one ulp below it and nothing may be lost."""
    from coldrec_b200 import ops
    U = np.zeros((300, 64), np.float32); I = np.zeros((100000, 64), np.float32)
    s, i, _ = ops.score_topk(cu(U), cu(I), 20, precision=ops.SCORE_TF32_CHECKED)
    assert (s.cpu().numpy() == 0).all()
    assert (i.cpu().numpy() == np.arange(20)[None, :]).all(), "ties resolve to the smallest ids"


def test_topk_merge_and_item_sharding_equals_single_sweep():
    """Item-sharded scoring + merge (what each GPU does before/after the candidate allgather) must give
    the single-sweep lists bit for bit (ids) — the comparator (score desc, id asc) is order independent."""
    from coldrec_b200 import ops
    U, I, uids, rowptr, col, flags = _synthetic_scoring_case(99, 400, 9000, 256, 64, 50, 0.05)
    s0, i0, _ = ops.score_topk(cu(U), cu(I), 20, user_ids=cu(uids), mask_rowptr=cu(rowptr), mask_col=cu(col), precision=0)
    parts_s, parts_i = [], []
    bounds = [0, 2000, 2500, 6111, 9000]
    for a, b in zip(bounds, bounds[1:]):
        s, i, _ = ops.score_topk(cu(U), cu(I[a:b]), 20, user_ids=cu(uids), item_id_base=a, mask_rowptr=cu(rowptr), mask_col=cu(col),
                                 precision=0)
        parts_s.append(s); parts_i.append(i)
    s1, i1 = ops.topk_merge(torch.stack(parts_s), torch.stack(parts_i))
    assert torch.equal(i0, i1) and torch.equal(s0, s1)
    # shard smaller than K pads with (-inf, -1) and still merges
    s, i, _ = ops.score_topk(cu(U), cu(I[:7]), 20, user_ids=cu(uids), precision=0)
    assert (i[:, 7:] == -1).all() and torch.isinf(s[:, 7:]).all()


def test_gather_rows_and_errors():
    from coldrec_b200 import ops
    src = torch.randn(100, 64, device=DEV)
    ids = torch.randint(0, 100, (37,), device=DEV, dtype=torch.int32)
    assert torch.equal(ops.gather_rows(src, ids), src[ids.long()])
    with pytest.raises(ValueError):
        ops.gather_rows(src.double(), ids)                       # no silent casts
    with pytest.raises(ValueError):
        ops.score_topk(src, torch.randn(50, 32, device=DEV), 20)  # d mismatch
    with pytest.raises(ValueError):
        ops.score_topk(src, src, 65)                              # K > CR_MAX_K


# ---------------------------------------------------------------------------------------------- towers
def _state(g, prefix):
    return {k[len(prefix):]: cu(v) for k, v in g.items() if k.startswith(prefix)}


def test_towers_vs_reference_golden():
    from coldrec_b200 import towers
    g = load_golden("towers")
    U, V, C = cu(g["user_emb"]), cu(g["item_emb"]), cu(g["item_content"])
    u, v = towers.dropoutnet_encode(_state(g, "dn."), U, V, None, C)
    assert_normwise(u, g["dn_user_out"]); assert_normwise(v, g["dn_item_out"])
    u, v = towers.heater_encode(_state(g, "ht."), U, V, C, int(g["ht_n_expert"]), float(g["ht_n_dropout"]))
    assert_normwise(u, g["ht_user_out"]); assert_normwise(v, g["ht_item_out"])
    cold = cu(g["cold_idx"], torch.int32)
    out = towers.gar_generate(_state(g, "gar."), C, rows=cold)
    assert_normwise(out, g["gar_cold_out"])
    table = V.clone()
    towers.gar_generate(_state(g, "gar."), C, rows=cold, out=table)       # GAR.py:44-46 cold-row overwrite
    want = g["item_emb"].copy(); want[g["cold_idx"]] = g["gar_cold_out"]
    assert_normwise(table, want)
    assert_normwise(towers.aldi_tower(_state(g, "aldi_u."), U), g["aldi_user_out"])
    assert_normwise(towers.aldi_tower(_state(g, "aldi_i."), C, rows=cold), g["aldi_cold_item_out"])


# ---------------------------------------------------------------------------------------------- tcgen05 probe
@pytest.mark.parametrize("shape", [(300, 64, 0, 200), (1000, 64, 2738, 200), (257, 200, 0, 100), (129, 100, 0, 64), (640, 2738, 0, 5),
                                   (77, 24, 0, 128), (513, 2738, 0, 205), (128, 36, 300, 256), (1, 8, 0, 16)])
def test_tower_layer_tc_3xtf32_matches_fp64(shape):
    """cr_linear_act_tc_f32 (tcgen05, hi.hi + lo.hi + hi.lo) against an fp64 evaluation of the same layer: widths that are not
    multiples of the 32-float K chunk or of the 16-column MMA N, the two-segment K loop ([V | content]), ragged last row
    tile, folded BatchNorm + tanh, the split outputs handed to the next layer, and the row scatter."""
    from coldrec_b200 import ops
    n, d1, d2, n_out = shape
    rng = np.random.default_rng(sum(shape))
    X1 = rng.standard_normal((n, d1)).astype(np.float32)
    X2 = (rng.standard_normal((n, d2)) * (rng.random((n, d2)) < 0.05)).astype(np.float32) if d2 else None
    W = (rng.standard_normal((n_out, d1 + d2)) * 0.05).astype(np.float32)
    b = rng.standard_normal(n_out).astype(np.float32) * 0.1
    sc = (rng.random(n_out) + 0.5).astype(np.float32)
    sh = (rng.standard_normal(n_out) * 0.1).astype(np.float32)
    Xcat = np.concatenate([X1, X2], 1).astype(np.float64) if d2 else X1.astype(np.float64)
    pre = (Xcat @ W.astype(np.float64).T + b) * sc + sh
    ref = np.tanh(pre)
    y, sp = ops.linear_act_tc(ops.split_tf32(cu(X1)), ops.split_tf32(cu(W)), cu(b), X2=ops.split_tf32(cu(X2)) if d2 else None,
                              scale=cu(sc), shift=cu(sh), act="tanh", want_split=True)
    assert_normwise(y, ref.astype(np.float32), tol=5e-6)          # half of the 1e-5 parity budget (K = 2,738 dense rows: ~3e-6)
    # the split output is the same value, exactly: hi + lo == y, hi is TF32-representable, padding columns are zero
    hi, lo = sp.hi.cpu().numpy(), sp.lo.cpu().numpy()
    assert np.array_equal((hi.astype(np.float64) + lo)[:, :n_out].astype(np.float32), y.cpu().numpy())
    assert (hi.view(np.uint32) & 0x1FFF).max() == 0 and not hi[:, n_out:].any() and not lo[:, n_out:].any()
    # raw fp32 rows, split inside the kernel (one HBM read): the same MMAs on the same hi / lo values -> the same bits
    y_raw, _ = ops.linear_act_tc(cu(X1), ops.split_tf32(cu(W)), cu(b), X2=cu(X2) if d2 else None, scale=cu(sc), shift=cu(sh), act="tanh")
    assert torch.equal(y_raw, y), "in-kernel split differs from the pre-split operands"
    # no epilogue, scattered rows
    perm = rng.permutation(n + 3)[:n].astype(np.int32)
    out = torch.full((n + 3, n_out), 7.0, device=DEV)
    ops.linear_act_tc(ops.split_tf32(cu(X1)), ops.split_tf32(cu(W)), None, X2=ops.split_tf32(cu(X2)) if d2 else None, out=out, yrow=cu(perm))
    want = np.full((n + 3, n_out), 7.0)
    want[perm] = Xcat @ W.astype(np.float64).T
    assert_normwise(out, want.astype(np.float32), tol=2e-6)
    # and the SIMT kernel agrees to the parity tolerance
    y2 = ops.linear_act(cu(X1), cu(W), cu(b), X2=cu(X2) if d2 else None, scale=cu(sc), shift=cu(sh), act="tanh")
    assert_normwise(y2, ref.astype(np.float32), tol=1e-5)


def test_tc_raw_scores_are_tf32_products():
    """The tensor-core sweep must see <q, x> with at most TF32 operand error: |err| <= 2^-9 |q||x|."""
    from coldrec_b200 import ops
    rng = np.random.default_rng(5)
    U = (rng.standard_normal((300, 64)) * 0.3).astype(np.float32)
    I = (rng.standard_normal((1000, 64)) * 0.3).astype(np.float32)
    s, i, dbg = ops.debug_tc_tile(cu(U), cu(I))
    ref = (U[:256].astype(np.float64) @ I[:96].astype(np.float64).T)
    bound = 2.0 ** -9 * np.linalg.norm(U[:256], axis=1)[:, None] * np.linalg.norm(I[:96], axis=1)[None, :] + 1e-6
    err = np.abs(dbg.cpu().numpy() - ref)
    assert (err <= bound).all(), f"max err {err.max()} (bound {bound.max()}); tile is not a TF32 product"
    assert err.max() > 1e-7, "suspiciously exact: is the tensor-core path really running?"
    exact = torch.topk(t(U) @ t(I).T, 20, dim=1)
    assert np.allclose(s.cpu().numpy(), exact.values.numpy(), atol=1e-5)
    assert np.array_equal(i.cpu().numpy().astype(np.int64), exact.indices.numpy())


def test_cgrc_frozen_cold_and_fsgnn_propagation_vs_reference_golden():
    """CGRC's frozen-cold layer list (model/CGRC.py:76-93), its mean over layers and FSGNN's _lightgcn (model/FSGNN.py:433-442)."""
    import coldrec_b200 as cr
    g, c = load_golden("graph"), load_golden("cgrc")
    n_u, n_i = int(g["user_num"]), int(g["item_num"])
    adj = sp.csr_matrix((g["adj_data"], g["adj_indices"], g["adj_indptr"]), shape=(n_u + n_i, n_u + n_i))
    G = cr.CsrGraph.from_scipy(adj, DEV)
    U, X = t(g["E0_user"]).to(DEV), t(c["item_x"]).to(DEV)
    cold = t(c["cold_item_idx"]).to(DEV)
    layers = cr.propagate_frozen_cold(G, U, X, 3, cold)
    assert len(layers) == 4
    for k, h in enumerate(layers):
        ref = c[f"frozen_L{k}"]
        assert np.abs(h.cpu().numpy() - ref).max() <= 1e-5 * np.abs(ref).max()
        if k:
            assert np.array_equal(h.cpu().numpy()[n_u + c["cold_item_idx"]], c["item_x"][c["cold_item_idx"]])
    none = cr.propagate_frozen_cold(G, U, X, 2, torch.zeros(0, dtype=torch.int64, device=DEV))
    assert np.abs(none[-1].cpu().numpy() - c["frozen_nocold_L2"]).max() <= 1e-5 * np.abs(c["frozen_nocold_L2"]).max()
    for L, ku, ki in ((3, "mean_user", "mean_item"), (2, "fsgnn_user", "fsgnn_item")):
        zu, zi = cr.propagate(G, U, X, L)
        for got, ref in ((zu, c[ku]), (zi, c[ki])):
            assert np.abs(got.cpu().numpy() - ref).max() <= 1e-5 * np.abs(ref).max()


def test_copy_rows_gather_scatter_and_errors():
    from coldrec_b200 import ops
    rng = np.random.default_rng(5)
    src = torch.from_numpy(rng.standard_normal((50, 32)).astype(np.float32)).to(DEV)
    dst = torch.zeros((80, 32), device=DEV)
    si = torch.from_numpy(rng.choice(50, 20, replace=False).astype(np.int32)).to(DEV)
    di = torch.from_numpy(rng.choice(80, 20, replace=False).astype(np.int32)).to(DEV)
    ops.copy_rows(src, dst, src_ids=si, dst_ids=di)
    want = torch.zeros((80, 32), device=DEV)
    want[di.long()] = src[si.long()]
    assert torch.equal(dst, want)
    out = ops.copy_rows(src, torch.zeros((20, 32), device=DEV), src_ids=si)               # pure gather
    assert torch.equal(out, src[si.long()])
    out = ops.copy_rows(src[:20].contiguous(), torch.zeros((80, 32), device=DEV), dst_ids=di)   # pure scatter
    assert torch.equal(out[di.long()], src[:20])
    with pytest.raises(ValueError):
        ops.copy_rows(src, dst, src_ids=si, dst_ids=di[:5].contiguous())
    with pytest.raises(ValueError):
        ops.copy_rows(src, torch.zeros((80, 16), device=DEV))
    assert ops.copy_rows(src, dst, src_ids=si[:0].contiguous(), dst_ids=di[:0].contiguous()) is dst


def test_host_batch_evaluator_stages_group_and_prefetches():
    """HostBatchEvaluator (the e2e arm of bench.py): pinned host plans in, top-K + metric strings out; the plan of step k+1 is
    copied on a copy stream while step k is swept.  Every step must equal ranking the same plan from device memory."""
    from coldrec_b200 import ops
    from coldrec_b200.dist import GridShardedFullRankScorer
    from coldrec_b200.scoring import EvalPlan, HostBatchEvaluator
    U, I, _, _, _, _ = _synthetic_scoring_case(77, 3000, 20000, 10, 64, 10)
    Ud, Id = cu(U), cu(I)
    rng = np.random.default_rng(8)
    n_q = 700
    plans = []
    for step in range(5):
        uids = rng.choice(3000, n_q, replace=False).astype(np.int32)
        rows = [np.sort(rng.choice(20000, int(rng.integers(0, 40)), replace=False)) for _ in range(n_q)]
        rp = np.zeros(n_q + 1, dtype=np.int64); np.cumsum([len(r) for r in rows], out=rp[1:])
        gt = [np.sort(rng.choice(20000, 4, replace=False)) for _ in range(n_q)]
        plans.append(dict(user_ids=torch.from_numpy(uids), mask_rowptr=torch.from_numpy(rp), mask_col=torch.from_numpy(np.concatenate(rows).astype(np.int32)),
                          gt_rowptr=torch.arange(0, 4 * n_q + 1, 4, dtype=torch.int64), gt_col=torch.from_numpy(np.concatenate(gt).astype(np.int32))))
    sc = GridShardedFullRankScorer(20, 1, ops.SCORE_TF32_CHECKED)
    hb = HostBatchEvaluator(sc, [10, 20], n_q, n_q * 40, n_q * 4, torch.device(DEV))
    pinned = [hb.pin(p) for p in plans]
    for k, hp in enumerate(pinned):
        perf = hb.run(Ud, Id, 0, hp, pinned[k + 1] if k + 1 < len(pinned) else None)
        torch.cuda.synchronize()
        dplan = EvalPlan.from_arrays(**{n: cu(v) for n, v in plans[k].items()})
        s, i = sc.topk(Ud, Id, 0, dplan)
        assert np.array_equal(hb.out_ids.numpy(), i.cpu().numpy()) and np.array_equal(hb.out_scores.numpy(), s.cpu().numpy()), f"step {k}"
        assert perf == sc.metrics(i, dplan, [10, 20], rounded=True)
        assert hb.h2d_bytes == sum(v.numel() * v.element_size() for v in hp.values())
