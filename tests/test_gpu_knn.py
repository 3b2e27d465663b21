"""-m gpu: content kNN through the fused scorer (wide tables, d up to 2740) and the kNN cold-row generator vs the oracle."""
import numpy as np
import pytest
import torch

from oracle import coldrec_oracle as O
from tests.helpers import builder_args, load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _check(ref_s, ref_i, got_s, got_i, sim_row):
    O.check_topk_parity(ref_s, ref_i.astype(np.int64), got_s, got_i.astype(np.int64), lambda j, ids: [float(sim_row(j)[int(x)]) for x in ids])


@pytest.mark.parametrize("shape", [(300, 2000, 300, 10), (257, 1500, 2738, 5), (100, 900, 516, 20), (64, 700, 260, 8)])
def test_knn_inner_product_wide_tables_vs_oracle(shape):
    from coldrec_b200 import knn
    n_q, n_v, d, k = shape
    rng = np.random.default_rng(d)
    Q = (rng.standard_normal((n_q, d)) / np.sqrt(d)).astype(np.float32)
    V = (rng.standard_normal((n_v, d)) / np.sqrt(d)).astype(np.float32)
    V[7] = V[3]                                          # an exact tie
    ref_s, ref_i = O.knn_inner_product(Q, V, k)
    s, i = knn.knn_inner_product(Q, V, k, DEV)
    sim = Q @ V.T
    _check(ref_s, ref_i, s.cpu().numpy(), i.cpu().numpy(), lambda j: sim[j])
    assert (np.diff(s.cpu().numpy(), axis=1) <= 0).all()


def test_cosine_knn_graph_excludes_self():
    from coldrec_b200 import knn
    rng = np.random.default_rng(1)
    F = rng.standard_normal((400, 300)).astype(np.float32)
    x = F.astype(np.float64); x = (x / np.maximum(np.linalg.norm(x, axis=1, keepdims=True), 1e-12)).astype(np.float32)
    ref_s, ref_i = O.knn_inner_product(x, x, 10, exclude_self=True)
    s, i = knn.cosine_knn_graph(F, 10, DEV)
    ii = i.cpu().numpy()
    assert not (ii == np.arange(400)[:, None]).any()
    sim = x @ x.T
    _check(ref_s, ref_i, s.cpu().numpy(), ii, lambda j: sim[j])


def test_knn_model_neighbours_and_generator_on_reference_data():
    """KNN._precompute_knn_neighbors + knn_search (model/KNN.py:63-88) on the reference-built dataset of the golden files."""
    from coldrec_b200 import knn
    g = load_golden("eval_item")
    data = O.OracleData(*builder_args(g))
    ref_nb = O.precompute_knn_neighbors(data, "item", 5)
    nb = knn.precompute_knn_neighbors(data, "item", 5, DEV)
    content = np.asarray(data.mapped_item_content, dtype=np.float32)
    sim = content[data.mapped_cold_item_idx] @ content.T
    same = (nb == ref_nb)
    for r, c in zip(*np.nonzero(~same)):                 # only genuine near-ties may differ
        assert abs(sim[r, nb[r, c]] - sim[r, ref_nb[r, c]]) <= 1e-5 * max(1.0, abs(sim[r, ref_nb[r, c]]))
    emb = torch.from_numpy(g["item_emb"])
    ref = O.knn_generate(emb, ref_nb)
    out = knn.knn_generate(emb.to(DEV), torch.from_numpy(ref_nb).to(DEV))
    assert (out.cpu() - ref).abs().max().item() <= 1e-5 * ref.abs().max().item()
