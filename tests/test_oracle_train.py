"""CPU: the training-step / sampler oracle vs. the golden vectors produced by the reference's own training loop body
(tests/golden/train.npz: model/LightGCN.py:21-28, model/MF.py:19-27, util/utils.py:123-157 run unmodified)."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

from oracle import coldrec_oracle as O
from tests.helpers import load_golden


@pytest.fixture(scope="module", autouse=True)
def _single_thread():
    n = torch.get_num_threads()
    torch.set_num_threads(1)
    yield
    torch.set_num_threads(n)


def _adj(g, gg):
    return sp.csr_matrix((gg["adj_data"], gg["adj_indices"], gg["adj_indptr"]), shape=(len(gg["adj_indptr"]) - 1,) * 2)


@pytest.mark.parametrize("tag,layers", [("lgcn", 3), ("mf", 0)])
def test_train_oracle_reproduces_reference_steps(tag, layers):
    g, gg = load_golden("train"), load_golden("graph")
    orc = O.TrainOracle(_adj(g, gg) if layers else None, g[f"{tag}_E0_user"], g[f"{tag}_E0_item"], layers, float(g["lr"]), float(g["reg"]))
    for s in range(int(g["n_steps"])):
        losses = orc.step(g[f"batch{s}_u"], g[f"batch{s}_i"], g[f"batch{s}_j"])
        assert np.allclose(losses, g[f"{tag}_loss{s}"], rtol=1e-6, atol=1e-7)
        for name, p in (("user", orc.user), ("item", orc.item)):
            assert np.allclose(p.grad.numpy(), g[f"{tag}_grad{s}_{name}"], rtol=1e-5, atol=1e-9)
            assert np.allclose(p.detach().numpy(), g[f"{tag}_param{s}_{name}"], rtol=0, atol=1e-7)
    for name, p in (("user", orc.user), ("item", orc.item)):
        st = orc.opt.state[p]
        assert np.allclose(st["exp_avg"].numpy(), g[f"{tag}_exp_avg_{name}"], rtol=1e-5, atol=1e-10)
        assert np.allclose(st["exp_avg_sq"].numpy(), g[f"{tag}_exp_avg_sq_{name}"], rtol=1e-5, atol=1e-14)


def test_batch_grads_and_adam_pieces_match_reference():
    g = load_golden("train")
    losses, gu, gi = O.bpr_batch_grads(torch.from_numpy(g["mf_E0_user"]), torch.from_numpy(g["mf_E0_item"]), g["batch0_u"], g["batch0_i"],
                                       g["batch0_j"], float(g["reg"]))
    assert np.allclose(losses, g["mf_loss0"], rtol=1e-6)
    assert np.allclose(gu.numpy(), g["mf_grad0_user"], rtol=1e-5, atol=1e-9)
    assert np.allclose(gi.numpy(), g["mf_grad0_item"], rtol=1e-5, atol=1e-9)
    n = int(g["n_steps"])
    p, m, v = O.adam_reference(g["mf_E0_user"], [g[f"mf_grad{s}_user"] for s in range(n)], float(g["lr"]))
    assert np.allclose(p, g[f"mf_param{n - 1}_user"], rtol=0, atol=1e-7)
    assert np.allclose(m, g["mf_exp_avg_user"], rtol=1e-6, atol=1e-12)
    assert np.allclose(v, g["mf_exp_avg_sq_user"], rtol=1e-6, atol=1e-16)


def test_reference_sampler_epoch_satisfies_the_pinned_invariants():
    g = load_golden("train")
    O.check_sampler_epoch(g["epoch_u"], g["epoch_i"], g["epoch_j"], g["train_u"], g["train_i"], int(g["n_item_table"]))
    for s in range(int(g["n_steps"])):           # the golden batches are the first batches of that epoch
        b = int(g["bs"])
        assert np.array_equal(g[f"batch{s}_u"], g["epoch_u"][s * b:(s + 1) * b])


def _train_csr(tu, ti, n_users, n_items):
    m = sp.csr_matrix((np.ones(len(tu)), (tu, ti)), shape=(n_users, n_items))
    m.sum_duplicates(); m.sort_indices()
    return m.indptr.astype(np.int64), m.indices.astype(np.int32)


def test_oracle_sampler_invariants_and_reproducibility():
    g = load_golden("train")
    tu, ti, n_items = g["train_u"], g["train_i"], int(g["n_item_table"])
    rp, col = _train_csr(tu, ti, int(tu.max()) + 1, n_items)
    u, i, j = O.sample_pairwise(tu, ti, rp, col, n_items, 2024, 3, 0, len(tu))
    O.check_sampler_epoch(u, i, j, tu, ti, n_items)
    # batches are slices of the epoch whatever the batch size; epochs and seeds differ
    u2, i2, j2 = O.sample_pairwise(tu, ti, rp, col, n_items, 2024, 3, 100, 57)
    assert np.array_equal(u2, u[100:157]) and np.array_equal(i2, i[100:157]) and np.array_equal(j2, j[100:157])
    u3, _, j3 = O.sample_pairwise(tu, ti, rp, col, n_items, 2024, 4, 0, len(tu))
    assert not np.array_equal(u3, u) and not np.array_equal(j3, j)
    u4, _, _ = O.sample_pairwise(tu, ti, rp, col, n_items, 7, 3, 0, len(tu))
    assert not np.array_equal(u4, u)


@pytest.mark.parametrize("n", [1, 2, 3, 17, 256, 257, 5000])
def test_feistel_permutation_is_a_bijection(n):
    p = O.feistel_perm(np.arange(n), n, 0xDEADBEEF, 0x12345)
    assert np.array_equal(np.sort(p), np.arange(n))
    if n >= 256:        # and it actually shuffles
        assert (p == np.arange(n)).mean() < 0.05


def test_oracle_negatives_are_uniform_over_non_train_items():
    rng = np.random.default_rng(5)
    n_users, n_items = 50, 40
    tu = np.repeat(np.arange(n_users), 10)
    ti = np.concatenate([rng.choice(n_items, 10, replace=False) for _ in range(n_users)])
    rp, col = _train_csr(tu, ti, n_users, n_items)
    counts = np.zeros((n_users, n_items))
    for e in range(40):
        u, _, j = O.sample_pairwise(tu, ti, rp, col, n_items, 99, e, 0, len(tu))
        np.add.at(counts, (u, j), 1)
    train = np.zeros((n_users, n_items), bool); train[tu, ti] = True
    assert counts[train].sum() == 0
    free = counts[~train].reshape(n_users, n_items - 10)        # 400 draws per user over 30 free items
    chi2 = (((free - 400 / 30) ** 2) / (400 / 30)).sum()
    dof = n_users * 29
    assert abs(chi2 - dof) < 6 * np.sqrt(2 * dof), chi2
