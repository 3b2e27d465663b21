"""BASELINE.json configs[0..2] at their full (dataset-shaped) sizes: CUDA path vs the CPU oracle on seeded synthetic data.

  C1  BPR-MF, MovieLens-shaped: 6,040 users x 3,706 items, ~1M interactions, item cold-start (data/README.md:8)
  C2  LightGCN 3-layer propagation + full ranking, CiteULike-shaped: 5,551 x 16,980, ~205k interactions (:10)
  C3  DropoutNet / Heater cold-item generation from 2,738-d content + cold-item top-20, XING-shaped: 106,881 x 20,519 (:11)

The split mimics data/split.py (80/20 warm/cold items, warm 8:1:1) with arrays instead of CSV files; eval users and
masks are built exactly as `_get_eval_cache` does (train items masked; 'cold' setting masks warm items).
"""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

from oracle import coldrec_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def cu(x, dtype=None):
    x = torch.from_numpy(np.ascontiguousarray(x)) if isinstance(x, np.ndarray) else x
    return (x.to(dtype) if dtype is not None else x).to(DEV)


def _interactions(rng, n_users, n_items, n_inter):
    pu = rng.lognormal(0.0, 1.0, n_users); pu /= pu.sum()
    pi = 1.0 / np.arange(1, n_items + 1) ** 1.0; pi = rng.permutation(pi); pi /= pi.sum()
    u = rng.choice(n_users, int(n_inter * 1.15), p=pu)
    i = rng.choice(n_items, int(n_inter * 1.15), p=pi)
    key = np.unique(u.astype(np.int64) * n_items + i)[:n_inter]
    key = key[rng.permutation(len(key))]
    return (key // n_items).astype(np.int64), (key % n_items).astype(np.int64)


def _split(rng, u, i, n_items):
    """warm/cold items 80/20; warm interactions 8:1:1 train/val/test; cold items' interactions -> cold test."""
    cold_item = np.zeros(n_items, dtype=bool)
    cold_item[rng.permutation(n_items)[:n_items // 5]] = True
    is_cold = cold_item[i]
    r = rng.random(len(u))
    train = ~is_cold & (r < 0.8)
    warm_test = ~is_cold & (r >= 0.9)
    return cold_item, train, warm_test, is_cold


def _csr_rows(users, items, sel_users, n_items):
    m = sp.csr_matrix((np.ones(len(users), np.int8), (users, items)), shape=(int(users.max(initial=0)) + 1 if len(users) else 1, n_items))
    m.sum_duplicates(); m.sort_indices()
    rows = [m.indices[m.indptr[x]:m.indptr[x + 1]] if x < m.shape[0] else np.zeros(0, np.int32) for x in sel_users]
    rowptr = np.zeros(len(sel_users) + 1, dtype=np.int64); np.cumsum([len(r) for r in rows], out=rowptr[1:])
    col = np.concatenate(rows).astype(np.int32) if rowptr[-1] else np.zeros(0, np.int32)
    return rowptr, col


def _check_eval(U, I, uids, mask_rp, mask_col, gt_rp, gt_col, flags, excl, precision):
    from coldrec_b200 import ops
    from coldrec_b200.evaluator import device_metrics
    col_mask = None if excl == 0 else np.nonzero(flags & excl)[0]
    ref_s, ref_i = O.evaluate_topk_dense(O.score_mf(torch.from_numpy(U), torch.from_numpy(I)), uids, mask_rp, mask_col.astype(np.int64),
                                         col_mask, 20, 4096)
    s, i, _ = ops.score_topk(cu(U), cu(I), 20, user_ids=cu(uids.astype(np.int32)), mask_rowptr=cu(mask_rp), mask_col=cu(mask_col),
                             item_flags=cu(flags) if excl else None, flag_exclude=excl, precision=precision)
    Ut, It = torch.from_numpy(U), torch.from_numpy(I)
    cm = set() if col_mask is None else set(col_mask.tolist())

    def exact(j, ids):
        masked = set(mask_col[mask_rp[j]:mask_rp[j + 1]].tolist())
        row = (Ut[uids[j]] @ It.T).numpy()
        return [O.MASK_SENTINEL if (int(x) in masked or int(x) in cm) else float(row[int(x)]) for x in ids]
    ties = O.check_topk_parity(ref_s, ref_i, s.cpu().numpy(), i.cpu().numpy().astype(np.int64), exact)
    got = device_metrics(i, cu(gt_rp), cu(gt_col), [10, 20], rounded=False)
    want = O.metrics_from_topk(ref_i, gt_rp, gt_col.astype(np.int64), [10, 20])
    if ties == 0:
        assert np.allclose(got, want, atol=1e-6, rtol=0), (got, want)
    return got


@pytest.mark.parametrize("precision", [0, 1], ids=["exact", "tf32"])
def test_c1_bprmf_movielens_shaped(precision):
    rng = np.random.default_rng(2024)
    n_users, n_items = 6040, 3706
    u, i = _interactions(rng, n_users, n_items, 1_000_209)
    cold_item, train, warm_test, is_cold = _split(rng, u, i, n_items)
    U = (rng.standard_normal((n_users, 64)) * 0.1).astype(np.float32)
    I = (rng.standard_normal((n_items, 64)) * 0.1).astype(np.float32)
    flags = np.where(cold_item, 1, 2).astype(np.uint8)
    for setting, sel, excl in (("warm", warm_test, 1), ("cold", is_cold, 2), ("all", warm_test | is_cold, 0)):
        eval_users = np.unique(u[sel])
        mask_rp, mask_col = _csr_rows(u[train], i[train], eval_users, n_items)
        gt_rp, gt_col = _csr_rows(u[sel], i[sel], eval_users, n_items)
        perf = _check_eval(U, I, eval_users, mask_rp, mask_col, gt_rp, gt_col, flags, excl, precision)
        assert 0.0 <= perf[1][3] <= 1.0


def test_c2_lightgcn_citeulike_shaped():
    from coldrec_b200 import CsrGraph, bipartite_norm_csr, propagate
    rng = np.random.default_rng(3)
    n_users, n_items = 5551, 16980
    u, i = _interactions(rng, n_users, n_items, 204_986)
    cold_item, train, warm_test, is_cold = _split(rng, u, i, n_items)
    bound = (6.0 / (n_users + 64)) ** 0.5
    E0u = ((rng.random((n_users, 64)) * 2 - 1) * bound).astype(np.float32)       # xavier_uniform, LightGCN.py:79-83
    E0i = ((rng.random((n_items, 64)) * 2 - 1) * (6.0 / (n_items + 64)) ** 0.5).astype(np.float32)
    adj = O.normalize_graph_mat(O.bipartite_adjacency(u[train], i[train], n_users, n_items))
    ref_u, ref_i = O.propagate(adj, torch.from_numpy(E0u), torch.from_numpy(E0i), 3)
    for G in (CsrGraph.from_scipy(adj, DEV), bipartite_norm_csr(cu(u[train]), cu(i[train]), n_users, n_items)):
        pu, pi = propagate(G, cu(E0u), cu(E0i), 3)
        for got, ref in ((pu, ref_u), (pi, ref_i)):
            assert (got.cpu() - ref).abs().max().item() <= 1e-5 * ref.abs().max().item()
    # full ranking on the propagated tables (the tables the oracle produced, so both sides score identical inputs)
    flags = np.where(cold_item, 1, 2).astype(np.uint8)
    eval_users = np.unique(u[warm_test | is_cold])
    mask_rp, mask_col = _csr_rows(u[train], i[train], eval_users, n_items)
    gt_rp, gt_col = _csr_rows(u[warm_test | is_cold], i[warm_test | is_cold], eval_users, n_items)
    _check_eval(ref_u.numpy() * 30, ref_i.numpy() * 30, eval_users, mask_rp, mask_col, gt_rp, gt_col, flags, 0, 1)


def test_c3_content_generators_xing_shaped():
    from coldrec_b200 import ops, towers
    rng = np.random.default_rng(4)
    n_users, n_items, C = 106_881, 20_519, 2738
    g = torch.Generator().manual_seed(4)
    U = torch.randn(n_users, 64, generator=g) * 0.1
    V = torch.randn(n_items, 64, generator=g) * 0.1
    content = (torch.rand(n_items, C, generator=g) < 0.02).float() * torch.randn(n_items, C, generator=g)   # sparse-ish like XING
    # DropoutNet towers (model/DropoutNet.py:155-213): [V|content] -> 200 -> 100 -> 64 with eval BatchNorm
    def lin(o, k, std):
        return torch.randn(o, k, generator=g) * std, torch.randn(o, generator=g) * 0.05
    def bn(n):
        return (torch.rand(n, generator=g) + 0.5, torch.randn(n, generator=g) * 0.1, torch.randn(n, generator=g) * 0.05,
                torch.rand(n, generator=g) * 0.5 + 0.5)
    sd = {}
    blocks = {"u": [], "v": []}
    for side, k0 in (("u", 64), ("v", 64 + C)):
        dims = [k0, 200, 100]
        for l in range(2):
            w, b = lin(dims[l + 1], dims[l], 0.08 if dims[l] < 1000 else 0.03)
            gam, bet, mu, var = bn(dims[l + 1])
            blocks[side].append((w, b, gam, bet, mu, var))
            sd.update({f"{side}_layers.{l}.layer.weight": w, f"{side}_layers.{l}.layer.bias": b, f"{side}_layers.{l}.bn.weight": gam,
                       f"{side}_layers.{l}.bn.bias": bet, f"{side}_layers.{l}.bn.running_mean": mu, f"{side}_layers.{l}.bn.running_var": var})
        w, b = lin(64, 100, 0.1)
        sd.update({f"{side}_emb.weight": w, f"{side}_emb.bias": b})
    ref_u, ref_v = O.dropoutnet_encode(U, V, None, content, blocks["u"], (sd["u_emb.weight"], sd["u_emb.bias"]), blocks["v"],
                                       (sd["v_emb.weight"], sd["v_emb.bias"]))
    got_u, got_v = towers.dropoutnet_encode({k: cu(v) for k, v in sd.items()}, cu(U), cu(V), None, cu(content))
    for got, ref in ((got_u, ref_u), (got_v, ref_v)):
        assert (got.cpu() - ref).abs().max().item() <= 1e-5 * ref.abs().max().item()
    # Heater (model/Heater.py:170-223): gate + shared expert MLP + blend + out_linear + final_trans
    hp = dict(gate_w=lin(5, C, 0.03), fc1=lin(200, C, 0.03), fc2=lin(64, 200, 0.08), out=lin(64, 64, 0.1), fin=lin(64, 64, 0.1))
    p = dict(gate_w=hp["gate_w"][0], gate_b=hp["gate_w"][1], fc1_w=hp["fc1"][0], fc1_b=hp["fc1"][1], fc2_w=hp["fc2"][0], fc2_b=hp["fc2"][1],
             out_w=hp["out"][0], out_b=hp["out"][1], fin_w=hp["fin"][0], fin_b=hp["fin"][1])
    ref_hu, ref_hv = O.heater_encode(U, V, content, p, 5, 0.5)
    hsd = {"gate.linear.weight": p["gate_w"], "gate.linear.bias": p["gate_b"], "fc.linear1.weight": p["fc1_w"], "fc.linear1.bias": p["fc1_b"],
           "fc.linear2.weight": p["fc2_w"], "fc.linear2.bias": p["fc2_b"], "out_linear.weight": p["out_w"], "out_linear.bias": p["out_b"],
           "final_trans.weight": p["fin_w"], "final_trans.bias": p["fin_b"]}
    got_hu, got_hv = towers.heater_encode({k: cu(v) for k, v in hsd.items()}, cu(U), cu(V), cu(content), 5, 0.5)
    for got, ref in ((got_hu, ref_hu), (got_hv, ref_hv)):
        assert (got.cpu() - ref).abs().max().item() <= 1e-5 * ref.abs().max().item()
    # cold-item top-20 on the generated tables for a slice of the users ('cold' setting: warm items masked)
    cold_item = np.zeros(n_items, dtype=bool); cold_item[rng.permutation(n_items)[:n_items // 5]] = True
    flags = np.where(cold_item, 1, 2).astype(np.uint8)
    eval_users = np.sort(rng.choice(n_users, 4096, replace=False))
    mask_rows = [np.sort(rng.choice(n_items, int(rng.integers(5, 80)), replace=False)) for _ in eval_users]
    mask_rp = np.zeros(len(eval_users) + 1, dtype=np.int64); np.cumsum([len(r) for r in mask_rows], out=mask_rp[1:])
    mask_col = np.concatenate(mask_rows).astype(np.int32)
    cold_ids = np.nonzero(cold_item)[0]
    gt_rows = [np.sort(rng.choice(cold_ids, 5, replace=False)) for _ in eval_users]
    gt_rp = np.arange(0, 5 * len(eval_users) + 1, 5, dtype=np.int64)
    gt_col = np.concatenate(gt_rows).astype(np.int32)
    _check_eval(ref_u.numpy(), ref_v.numpy(), eval_users, mask_rp, mask_col, gt_rp, gt_col, flags, 2, 1)
