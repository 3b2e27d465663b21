"""-m gpu: the training-side kernels (K5 fused BPR step + Adam, K6 pairwise sampler) through the C ABI vs. the oracle
and vs. the golden vectors the reference's own training loop produced (tests/golden/train.npz).

Tolerances: losses 1e-6 relative; gradients and propagated tables norm-wise 1e-5 (max|d| <= 1e-5 max|ref|, the
north-star bound for propagated embeddings); Adam update of a given gradient 1e-6 relative; sampler output bit-exact
against the oracle's restatement of the same counter-based generator."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

from oracle import coldrec_oracle as O
from tests.helpers import load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def cu(x, dtype=None):
    x = torch.from_numpy(np.ascontiguousarray(x)) if isinstance(x, np.ndarray) else x
    return (x.to(dtype) if dtype is not None else x).to(DEV)


def normwise(got, ref, tol=1e-5):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    err, scale = np.abs(got - ref).max(), np.abs(ref).max()
    assert err <= tol * scale + 1e-30, f"max|d|={err:.3e} vs {tol}*max|ref|={tol * scale:.3e}"


def _adj(gg):
    return sp.csr_matrix((gg["adj_data"], gg["adj_indices"], gg["adj_indptr"]), shape=(len(gg["adj_indptr"]) - 1,) * 2)


def test_bpr_fwd_bwd_matches_reference_mf_step():
    from coldrec_b200 import ops
    g = load_golden("train")
    U, I = cu(g["mf_E0_user"]), cu(g["mf_E0_item"])
    gu, gi = torch.zeros_like(U), torch.zeros_like(I)
    loss = ops.bpr_fwd_bwd(U, I, cu(g["batch0_u"], torch.int32), cu(g["batch0_i"], torch.int32), cu(g["batch0_j"], torch.int32),
                           float(g["reg"]), gu, gi)
    assert np.allclose(loss.cpu().numpy()[:3], g["mf_loss0"], rtol=1e-6)
    normwise(gu.cpu().numpy(), g["mf_grad0_user"])
    normwise(gi.cpu().numpy(), g["mf_grad0_item"])


@pytest.mark.parametrize("d", [32, 64, 96, 128, 200, 256])
@pytest.mark.parametrize("B", [1, 37, 4096])
def test_bpr_fwd_bwd_vs_autograd(d, B):
    from coldrec_b200 import ops
    rng = np.random.default_rng(d * 7 + B)
    n_users, n_items = 300, 500
    U = (rng.standard_normal((n_users, d)) * 0.3).astype(np.float32)
    I = (rng.standard_normal((n_items, d)) * 0.3).astype(np.float32)
    u, i, j = rng.integers(0, n_users, B), rng.integers(0, n_items, B), rng.integers(0, n_items, B)   # duplicates accumulate
    losses, ref_gu, ref_gi = O.bpr_batch_grads(torch.from_numpy(U), torch.from_numpy(I), u, i, j, 1e-2)
    gu, gi = torch.zeros((n_users, d), device=DEV), torch.zeros((n_items, d), device=DEV)
    loss = ops.bpr_fwd_bwd(cu(U), cu(I), cu(u, torch.int32), cu(i, torch.int32), cu(j, torch.int32), 1e-2, gu, gi)
    got = loss.cpu().numpy()
    assert np.allclose(got[:2], losses[:2], rtol=5e-6)
    # the reference's torch.norm sums B*d squares in fp32 (the kernel sums them in fp64): 1e-5 apart at 10^6 elements
    assert np.isclose(got[2], losses[2], rtol=5e-5)
    normwise(gu.cpu().numpy(), ref_gu.numpy())
    normwise(gi.cpu().numpy(), ref_gi.numpy())


def test_bpr_rejects_bad_arguments():
    from coldrec_b200 import ops
    U, I = torch.zeros((4, 64), device=DEV), torch.zeros((4, 64), device=DEV)
    idx = torch.zeros(3, dtype=torch.int32, device=DEV)
    with pytest.raises(ValueError):
        ops.bpr_fwd_bwd(U, I, idx, idx, idx[:2], 0.0, torch.zeros_like(U), torch.zeros_like(I))
    with pytest.raises(ValueError):
        ops.bpr_fwd_bwd(U, I, idx.to(torch.int64), idx, idx, 0.0, torch.zeros_like(U), torch.zeros_like(I))
    with pytest.raises(ValueError):
        ops.bpr_fwd_bwd(U.cpu(), I, idx, idx, idx, 0.0, torch.zeros_like(U), torch.zeros_like(I))


@pytest.mark.parametrize("n", [1, 7, 64 * 390, 1000003])
def test_adam_step_vs_torch_optim(n):
    from coldrec_b200 import ops
    rng = np.random.default_rng(n)
    p0 = (rng.standard_normal(n) * 0.1).astype(np.float32)
    grads = [(rng.standard_normal(n) * 10.0 ** rng.uniform(-6, 0)).astype(np.float32) for _ in range(4)]
    ref_p, ref_m, ref_v = O.adam_reference(p0, grads, 5e-3)
    p, m, v = cu(p0), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    for s, gr in enumerate(grads):
        ops.adam_step(p, cu(gr), m, v, s + 1, 5e-3)
    assert np.allclose(m.cpu().numpy(), ref_m, rtol=1e-6, atol=1e-12)
    assert np.allclose(v.cpu().numpy(), ref_v, rtol=1e-6, atol=1e-20)
    assert np.allclose(p.cpu().numpy(), ref_p, rtol=0, atol=2e-7)


@pytest.mark.parametrize("tag,layers", [("lgcn", 3), ("mf", 0)])
def test_train_step_reproduces_reference_training(tag, layers):
    """Three optimisation steps of the reference (LightGCN 3 layers / MF) on its own batches: losses, gradients w.r.t.
    the parameters, parameters and Adam state."""
    import coldrec_b200 as cr
    g, gg = load_golden("train"), load_golden("graph")
    graph = cr.CsrGraph.from_scipy(_adj(gg), DEV) if layers else None
    step = cr.BprTrainStep(graph, cu(g[f"{tag}_E0_user"]), cu(g[f"{tag}_E0_item"]), layers, float(g["lr"]), float(g["reg"]))
    n_u = step.n_users
    for s in range(int(g["n_steps"])):
        b = [cu(g[f"batch{s}_{k}"], torch.int32) for k in "uij"]
        grad = step.gradients(*b).clone()
        assert np.allclose(step.loss.cpu().numpy()[:3], g[f"{tag}_loss{s}"], rtol=2e-6)
        ref_g = np.concatenate([g[f"{tag}_grad{s}_user"], g[f"{tag}_grad{s}_item"]])
        normwise(grad.cpu().numpy(), ref_g)
        step.steps += 1
        cr.ops.adam_step(step.ego, grad, step.exp_avg, step.exp_avg_sq, step.steps, step.lr)
        ref_p = np.concatenate([g[f"{tag}_param{s}_user"], g[f"{tag}_param{s}_item"]])
        # Adam divides by sqrt(v): where |g| is within fp32 noise of zero the update direction is ill-conditioned
        # (step 1 moves every element by ~lr * sign(g)), so those few elements are excluded from the tight bound.
        solid = np.abs(ref_g) > 1e-4 * np.abs(ref_g).max()
        d = np.abs(step.ego.cpu().numpy() - ref_p)
        assert d[solid].max() <= 1e-5 * np.abs(ref_p).max(), d[solid].max()
        assert d.max() <= 2.5 * float(g["lr"])
        step.ego.copy_(cu(ref_p))       # continue from the reference's parameters: every step's gradient is then comparable at 1e-5
        if layers:                      # propagation spreads the gradient over (almost) every row; MF touches batch rows only
            assert solid.mean() > 0.8
    ref_m = np.concatenate([g[f"{tag}_exp_avg_user"], g[f"{tag}_exp_avg_item"]])
    # three chained steps: the few ill-conditioned elements above feed back into the later gradients
    normwise(step.exp_avg.cpu().numpy(), ref_m, tol=1e-4)
    assert step.user_emb.shape[0] == n_u


def test_train_step_fused_call_equals_the_pieces():
    import coldrec_b200 as cr
    g, gg = load_golden("train"), load_golden("graph")
    graph = cr.CsrGraph.from_scipy(_adj(gg), DEV)
    a = cr.BprTrainStep(graph, cu(g["lgcn_E0_user"]), cu(g["lgcn_E0_item"]), 3, float(g["lr"]), float(g["reg"]))
    b = [cu(g[f"batch0_{k}"], torch.int32) for k in "uij"]
    loss = a.step(*b)
    assert np.allclose(loss.cpu().numpy()[:3], g["lgcn_loss0"], rtol=2e-6)
    eu, ei = a.embeddings()
    orc = O.TrainOracle(_adj(gg), a.user_emb.cpu().numpy(), a.item_emb.cpu().numpy(), 3, 1e-3, 0.0)
    with torch.no_grad():
        ru, ri = orc.forward()
    normwise(eu.cpu().numpy(), ru.numpy()); normwise(ei.cpu().numpy(), ri.numpy())


def test_captured_train_step_equals_eager_steps():
    """The CUDA-graph replay (device-side Adam factors) walks the same trajectory as eager steps, 300 steps long so the
    pinned-slot ring wraps."""
    import coldrec_b200 as cr
    g, gg = load_golden("train"), load_golden("graph")
    graph = cr.CsrGraph.from_scipy(_adj(gg), DEV)
    mk = lambda: cr.BprTrainStep(graph, cu(g["lgcn_E0_user"]), cu(g["lgcn_E0_item"]), 3, 1e-3, float(g["reg"]))
    eager, fast = mk(), mk()
    cap = fast.capture(int(g["bs"]))
    smp = cr.PairwiseSampler(cu(g["train_u"]), cu(g["train_i"]), fast.n_users, int(g["n_item_table"]), seed=5)
    B = int(g["bs"])
    for k in range(300):
        begin = (k * B) % (smp.n_pairs - B)
        u, i, j = smp.batch(k // 8, begin, B)
        le = eager.step(u, i, j).clone()
        smp.batch(k // 8, begin, B, out=cap.idx)
        lf = cap()
        if k % 50 == 0 or k == 299:
            assert torch.allclose(le, lf, rtol=1e-4, atol=1e-6), (k, le, lf)
    assert fast.steps == eager.steps == 300
    # fp32 atomics order differs run to run; Adam's sqrt(v) normalisation keeps the two trajectories within a few lr
    assert (fast.ego - eager.ego).abs().max().item() <= 1e-2 * eager.ego.abs().max().item() + 5e-3
    normwise(fast.exp_avg_sq.cpu().numpy(), eager.exp_avg_sq.cpu().numpy(), tol=1e-2)
    # a batch passed as tensors
    u, i, j = smp.batch(99, 0, B)
    cap(u, i, j)
    with pytest.raises(ValueError):
        cap(u[:10], i[:10], j[:10])


def _train_csr(tu, ti, n_users, n_items):
    m = sp.csr_matrix((np.ones(len(tu)), (tu, ti)), shape=(n_users, n_items))
    m.sum_duplicates(); m.sort_indices()
    return m.indptr.astype(np.int64), m.indices.astype(np.int32)


def test_sampler_bit_exact_vs_oracle_on_reference_pairs():
    import coldrec_b200 as cr
    g = load_golden("train")
    tu, ti, n_items = g["train_u"], g["train_i"], int(g["n_item_table"])
    n_users = int(tu.max()) + 1
    rp, col = _train_csr(tu, ti, n_users, n_items)
    smp = cr.PairwiseSampler(cu(tu), cu(ti), n_users, n_items, seed=2024)
    assert np.array_equal(smp.train_rowptr.cpu().numpy(), rp) and np.array_equal(smp.train_col.cpu().numpy(), col)
    for epoch in (0, 1, 2 ** 33 + 5):
        ref = O.sample_pairwise(tu, ti, rp, col, n_items, 2024, epoch, 0, len(tu))
        got = [np.concatenate([x.cpu().numpy() for x in col_]) for col_ in zip(*smp.epoch(epoch, 256))]
        for a, b in zip(got, ref):
            assert np.array_equal(a.astype(np.int64), b)
        O.check_sampler_epoch(got[0], got[1], got[2], tu, ti, n_items)
    assert int(smp.n_exhausted.item()) == 0


def test_sampler_large_epoch_invariants_and_uniformity():
    import coldrec_b200 as cr
    rng = np.random.default_rng(8)
    n_users, n_items, per = 20000, 3000, 50
    tu = np.repeat(np.arange(n_users), per)
    ti = (np.argsort(rng.random((n_users, n_items)), axis=1)[:, :per]).reshape(-1)
    smp = cr.PairwiseSampler(cu(tu), cu(ti), n_users, n_items, seed=11)
    u, i, j = (x.cpu().numpy().astype(np.int64) for x in smp.batch(3, 0, smp.n_pairs))
    O.check_sampler_epoch(u, i, j, tu, ti, n_items)
    # a strided slice of the oracle agrees (positions are independent of the launch geometry)
    rp, col = _train_csr(tu, ti, n_users, n_items)
    ref = O.sample_pairwise(tu, ti, rp, col, n_items, 11, 3, 777777, 500)
    assert np.array_equal(u[777777:778277], ref[0]) and np.array_equal(j[777777:778277], ref[2])
    # negatives are uniform over the item table up to the (2 %) rejection of train items: chi-square over items
    cnt = np.bincount(j, minlength=n_items).astype(np.float64)
    expect = len(j) / n_items
    chi2 = ((cnt - expect) ** 2 / expect).sum()
    assert abs(chi2 - (n_items - 1)) < 8 * np.sqrt(2 * (n_items - 1)), chi2
    # the shuffled stream is not ordered by user
    assert np.abs(np.corrcoef(np.arange(len(u)), u)[0, 1]) < 0.01


def test_sampler_user_who_saw_everything_is_reported():
    import coldrec_b200 as cr
    n_items = 8
    tu = np.concatenate([np.zeros(n_items, np.int64), np.array([1])])
    ti = np.concatenate([np.arange(n_items), np.array([3])])
    smp = cr.PairwiseSampler(cu(tu), cu(ti), 2, n_items, seed=1)
    u, i, j = smp.batch(0, 0, smp.n_pairs)
    assert int(smp.n_exhausted.item()) == n_items            # user 0 has no valid negative (the reference would spin forever)
    jj = j.cpu().numpy()[u.cpu().numpy() == 1]
    assert len(jj) == 1 and jj[0] != 3
