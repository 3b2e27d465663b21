"""The CPU oracle restatement vs. golden vectors produced by the UNMODIFIED reference."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

from oracle import coldrec_oracle as O
from tests.helpers import builder_args, load_golden, rec_from_golden, t


@pytest.fixture(scope="module", autouse=True)
def _single_thread():
    n = torch.get_num_threads()
    torch.set_num_threads(1)
    yield
    torch.set_num_threads(n)


def _gt(data, typ, split="test"):
    return getattr(data, {"all": "overall", "warm": "warm", "cold": "cold"}[typ] + f"_{split}_set")


@pytest.mark.parametrize("name", ["eval_item", "eval_user", "eval_tiny"])
def test_id_maps_and_cold_idx(name):
    g = load_golden(name)
    data = O.OracleData(*builder_args(g))
    assert [data.id2user[i] for i in range(len(data.user))] == g["id2user"].tolist()
    assert [data.id2item[i] for i in range(len(data.item))] == g["id2item"].tolist()
    for k in ("mapped_cold_item_idx", "mapped_warm_item_idx", "mapped_cold_user_idx", "mapped_warm_user_idx"):
        assert np.array_equal(np.asarray(getattr(data, k)), g[k])


@pytest.mark.parametrize("name", ["eval_item", "eval_user", "eval_tiny"])
@pytest.mark.parametrize("typ", ["all", "cold", "warm"])
def test_evaluate_and_metrics_mf(name, typ):
    g = load_golden(name)
    data = O.OracleData(*builder_args(g))
    ue, ie = t(g["user_emb"]), t(g["item_emb"])
    gt = _gt(data, typ)
    rec = O.evaluate(data, O.batch_predict_mf(data, ue, ie), gt, typ, str(g["cold_object"]), 20, int(g["batch_size"]))
    p = f"mf_test_{typ}"
    assert list(rec.keys()) == g[f"{p}_users"].tolist()
    got_s = np.array([[s for _, s in rec[u]] for u in rec], dtype=np.float32)
    got_i = np.array([[i for i, _ in rec[u]] for u in rec], dtype=np.int64)
    assert np.array_equal(got_s, g[f"{p}_scores"])
    unmasked = g[f"{p}_scores"] > -1e8          # ids among masked (-1e9) ties are unspecified
    assert np.array_equal(got_i[unmasked], g[f"{p}_raw_ids"][unmasked])
    measure, perf = O.ranking_evaluation(gt, rec_from_golden(g, p), [10, 20])
    assert measure == g[f"{p}_measure"].tolist()
    assert np.array_equal(np.array(perf), g[f"{p}_performance"])
    unr = O.ranking_metrics_unrounded(gt, rec_from_golden(g, p), [10, 20])
    assert np.allclose(np.array(unr), g[f"{p}_performance"], atol=5.1e-6, rtol=0)


def test_valid_all_single_n():
    g = load_golden("eval_item")
    data = O.OracleData(*builder_args(g))
    measure, perf = O.ranking_evaluation(data.overall_valid_set, rec_from_golden(g, "mf_valid_all"), [20])
    assert measure == g["mf_valid_all_measure"].tolist()


def test_metrics_from_topk_matches_dict_form():
    g = load_golden("eval_item")
    data = O.OracleData(*builder_args(g))
    p = "mf_test_all"
    gt = data.overall_test_set
    users = g[f"{p}_users"].tolist()
    rowptr = np.zeros(len(users) + 1, dtype=np.int64)
    cols = []
    for j, u in enumerate(users):
        ids = sorted(data.item[i] for i in gt[u])
        cols += ids
        rowptr[j + 1] = len(cols)
    got = O.metrics_from_topk(g[f"{p}_dense_ids"], rowptr, np.array(cols), [10, 20])
    want = O.ranking_metrics_unrounded(gt, rec_from_golden(g, p), [10, 20])
    assert np.allclose(got, want, atol=1e-12, rtol=0)


@pytest.mark.parametrize("typ", ["all", "cold"])
def test_evaluate_aldi_dual_table(typ):
    g = load_golden("eval_item")
    data = O.OracleData(*builder_args(g))
    fn_dense = O.score_aldi(t(g["user_emb"]), t(g["aldi_cold_user_emb"]), t(g["item_emb"]),
                            data.mapped_warm_item_idx, data.mapped_cold_item_idx)
    fn = lambda users: fn_dense(torch.tensor(data.get_user_id_list(users)))
    gt = _gt(data, typ)
    rec = O.evaluate(data, fn, gt, typ, "item", 20, int(g["batch_size"]))
    p = f"aldi_test_{typ}"
    got_s = np.array([[s for _, s in rec[u]] for u in rec], dtype=np.float32)
    got_i = np.array([[i for i, _ in rec[u]] for u in rec], dtype=np.int64)
    assert np.array_equal(got_s, g[f"{p}_scores"])
    assert np.array_equal(got_i, g[f"{p}_raw_ids"])


def test_evaluate_vbpr_two_products():
    g = load_golden("eval_item")
    data = O.OracleData(*builder_args(g))
    fn_dense = O.score_vbpr(t(g["user_emb"]), t(g["item_emb"]), t(g["vbpr_user_aux"]), t(g["vbpr_item_aux"]))
    fn = lambda users: fn_dense(torch.tensor(data.get_user_id_list(users)))
    rec = O.evaluate(data, fn, data.overall_test_set, "all", "item", 20, int(g["batch_size"]))
    got_i = np.array([[i for i, _ in rec[u]] for u in rec], dtype=np.int64)
    got_s = np.array([[s for _, s in rec[u]] for u in rec], dtype=np.float32)
    assert np.array_equal(got_i, g["vbpr_test_all_raw_ids"])
    assert np.array_equal(got_s, g["vbpr_test_all_scores"])


def test_dense_eval_equals_dict_eval():
    g = load_golden("eval_item")
    data = O.OracleData(*builder_args(g))
    p = "mf_test_warm"
    users = g[f"{p}_users"].tolist()
    uids = data.get_user_id_list(users)
    rowptr = np.zeros(len(users) + 1, dtype=np.int64)
    cols = []
    for j, u in enumerate(users):
        cols += data.get_item_id_list(list(data.training_set_u[u].keys())).tolist()
        rowptr[j + 1] = len(cols)
    s, i = O.evaluate_topk_dense(O.score_mf(t(g["user_emb"]), t(g["item_emb"])), uids, rowptr, np.array(cols, dtype=np.int64),
                                 data.mapped_cold_item_idx, 20, int(g["batch_size"]))
    assert np.array_equal(s, g[f"{p}_scores"])
    assert np.array_equal(i, g[f"{p}_dense_ids"])


def test_adjacency_and_normalisation():
    g = load_golden("graph")
    ui = O.bipartite_adjacency(g["train_u"], g["train_i"], int(g["user_num"]), int(g["item_num"])).tocsr()
    ui.sort_indices()
    assert np.array_equal(ui.indptr, g["ui_indptr"]) and np.array_equal(ui.indices, g["ui_indices"])
    assert np.array_equal(ui.data, g["ui_data"])
    na = O.normalize_graph_mat(ui).tocsr()
    na.sort_indices()
    assert np.array_equal(na.indptr, g["adj_indptr"]) and np.array_equal(na.indices, g["adj_indices"])
    assert np.array_equal(na.data.astype(np.float32), g["adj_data"])


def _norm_adj(g):
    import scipy.sparse as sp
    n = int(g["user_num"]) + int(g["item_num"])
    return sp.csr_matrix((g["adj_data"], g["adj_indices"], g["adj_indptr"]), shape=(n, n))


@pytest.mark.parametrize("L", [1, 2, 3])
def test_lightgcn_propagate(L):
    g = load_golden("graph")
    u, i = O.propagate(_norm_adj(g), t(g["E0_user"]), t(g["E0_item"]), L)
    assert np.allclose(u.numpy(), g[f"lgcn_L{L}_user"], atol=1e-7, rtol=0)
    assert np.allclose(i.numpy(), g[f"lgcn_L{L}_item"], atol=1e-7, rtol=0)


def test_simgcl_and_ngcf_propagate():
    g = load_golden("graph")
    u, i = O.propagate(_norm_adj(g), t(g["E0_user"]), t(g["E0_item"]), 3, include_ego=False)
    assert np.allclose(u.numpy(), g["simgcl_L3_user"], atol=1e-7, rtol=0)
    assert np.allclose(i.numpy(), g["simgcl_L3_item"], atol=1e-7, rtol=0)
    Wgc = [(t(g[f"ngcf_Wgc{l}_w"]), t(g[f"ngcf_Wgc{l}_b"])) for l in range(2)]
    Wbi = [(t(g[f"ngcf_Wbi{l}_w"]), t(g[f"ngcf_Wbi{l}_b"])) for l in range(2)]
    u, i = O.propagate_ngcf(_norm_adj(g), t(g["E0_user"]), t(g["E0_item"]), Wgc, Wbi)
    assert np.allclose(u.numpy(), g["ngcf_L2_user"], atol=1e-6, rtol=0)
    assert np.allclose(i.numpy(), g["ngcf_L2_item"], atol=1e-6, rtol=0)


def _dn_blocks(g, side):
    blocks = []
    for l in range(2):
        p = f"dn.{side}_layers.{l}."
        blocks.append(tuple(t(g[p + k]) for k in ("layer.weight", "layer.bias", "bn.weight", "bn.bias",
                                                   "bn.running_mean", "bn.running_var")))
    return blocks, (t(g[f"dn.{side}_emb.weight"]), t(g[f"dn.{side}_emb.bias"]))


def test_towers():
    g = load_golden("towers")
    U, V, C = t(g["user_emb"]), t(g["item_emb"]), t(g["item_content"])
    ub, uo = _dn_blocks(g, "u")
    vb, vo = _dn_blocks(g, "v")
    u, v = O.dropoutnet_encode(U, V, None, C, ub, uo, vb, vo)
    assert np.allclose(u.numpy(), g["dn_user_out"], atol=1e-6, rtol=0)
    assert np.allclose(v.numpy(), g["dn_item_out"], atol=1e-6, rtol=0)

    p = dict(gate_w=t(g["ht.gate.linear.weight"]), gate_b=t(g["ht.gate.linear.bias"]),
             fc1_w=t(g["ht.fc.linear1.weight"]), fc1_b=t(g["ht.fc.linear1.bias"]),
             fc2_w=t(g["ht.fc.linear2.weight"]), fc2_b=t(g["ht.fc.linear2.bias"]),
             out_w=t(g["ht.out_linear.weight"]), out_b=t(g["ht.out_linear.bias"]),
             fin_w=t(g["ht.final_trans.weight"]), fin_b=t(g["ht.final_trans.bias"]))
    u, v = O.heater_encode(U, V, C, p, int(g["ht_n_expert"]), float(g["ht_n_dropout"]))
    assert np.allclose(u.numpy(), g["ht_user_out"], atol=1e-6, rtol=0)
    assert np.allclose(v.numpy(), g["ht_item_out"], atol=1e-6, rtol=0)

    cold = t(g["cold_idx"])
    out = O.gar_generate(C[cold], t(g["gar.0.weight"]), t(g["gar.0.bias"]), t(g["gar.2.weight"]), t(g["gar.2.bias"]))
    assert np.allclose(out.numpy(), g["gar_cold_out"], atol=1e-6, rtol=0)

    tower = lambda pre, x: O.aldi_tower(x, t(g[pre + "fc1.weight"]), t(g[pre + "fc1.bias"]), t(g[pre + "bn.weight"]),
                                        t(g[pre + "bn.bias"]), t(g[pre + "bn.running_mean"]), t(g[pre + "bn.running_var"]),
                                        t(g[pre + "fc2.weight"]), t(g[pre + "fc2.bias"]))
    assert np.allclose(tower("aldi_u.", U).numpy(), g["aldi_user_out"], atol=1e-6, rtol=0)
    assert np.allclose(tower("aldi_i.", C[cold]).numpy(), g["aldi_cold_item_out"], atol=1e-6, rtol=0)


def test_chunked_dense_eval_equals_unchunked():
    rng = np.random.default_rng(17)
    U = torch.from_numpy((rng.standard_normal((300, 64)) * 0.2).astype(np.float32))
    I = torch.from_numpy((rng.standard_normal((5000, 64)) * 0.2).astype(np.float32))
    uids = rng.choice(300, 150, replace=False)
    rows = [np.sort(rng.choice(5000, int(rng.integers(0, 60)), replace=False)) for _ in uids]
    rowptr = np.zeros(len(uids) + 1, dtype=np.int64); np.cumsum([len(r) for r in rows], out=rowptr[1:])
    col = np.concatenate(rows).astype(np.int64)
    cmask = rng.choice(5000, 1000, replace=False)
    for cm in (None, cmask):
        s0, i0 = O.evaluate_topk_dense(O.score_mf(U, I), uids, rowptr, col, cm, 20, 64)
        s1, i1 = O.evaluate_topk_dense_chunked(U, I, uids, rowptr, col, cm, 20, user_batch=37, item_chunk=777)
        assert np.allclose(s0, s1, atol=1e-6, rtol=0)
        assert (i0 == i1).mean() > 0.999


def test_cgrc_fsgnn_propagation_variants_vs_reference_golden():
    """model/CGRC.py:64-93 and model/FSGNN.py:433-442, outputs of the reference's own functions (oracle/make_golden.py --only cgrc)."""
    g, c = load_golden("graph"), load_golden("cgrc")
    n_u, n_i = int(g["user_num"]), int(g["item_num"])
    adj = sp.csr_matrix((g["adj_data"], g["adj_indices"], g["adj_indptr"]), shape=(n_u + n_i, n_u + n_i))
    U, X = t(g["E0_user"]), t(c["item_x"])
    layers = O.propagate_frozen_cold(adj, U, X, 3, c["cold_item_idx"])
    assert len(layers) == 4
    for k, h in enumerate(layers):
        assert np.abs(h.numpy() - c[f"frozen_L{k}"]).max() <= 1e-7
        if k:
            assert np.array_equal(h.numpy()[n_u + c["cold_item_idx"]], c["item_x"][c["cold_item_idx"]])
    none = O.propagate_frozen_cold(adj, U, X, 2, np.zeros(0, dtype=np.int64))
    assert np.abs(none[-1].numpy() - c["frozen_nocold_L2"]).max() <= 1e-7
    zu, zi = O.propagate(adj, U, X, 3)                     # CGRC._lightgcn_mean_all_layers == the LightGCN encoder
    assert np.abs(zu.numpy() - c["mean_user"]).max() <= 1e-7 and np.abs(zi.numpy() - c["mean_item"]).max() <= 1e-7
    fu, fi = O.propagate(adj, U, X, 2)                     # FSGNN._lightgcn: stack(dim=0).mean(dim=0), 2 layers
    assert np.abs(fu.numpy() - c["fsgnn_user"]).max() <= 1e-7 and np.abs(fi.numpy() - c["fsgnn_item"]).max() <= 1e-7
