"""-m gpu: the array-based data builder as the data object of the trainer API and as the producer of eval plans,
device graphs and samplers — same results as with the reference's builder (golden vectors)."""
import numpy as np
import pytest
import torch

from oracle import coldrec_oracle as O
from tests.helpers import builder_args, load_golden, t
from tests.test_gpu_parity import DEV, _exact_scores_fn, _trainer, cu

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["eval_item", "eval_user"])
@pytest.mark.parametrize("typ", ["all", "cold", "warm"])
def test_trainer_on_array_builder_vs_reference_golden(name, typ):
    from coldrec_b200 import ArrayDataBuilder, FullRankScorer, ops
    from coldrec_b200.evaluator import device_metrics
    g = load_golden(name)
    data = ArrayDataBuilder(*builder_args(g))
    ref_data = O.OracleData(*builder_args(g))
    cold_object = str(g["cold_object"])
    tr = _trainer(data, cold_object, ops.SCORE_TF32_CHECKED)
    tr.user_emb, tr.item_emb = cu(g["user_emb"]), cu(g["item_emb"])
    rec = tr.test(typ)                                   # the reference API, dict views built lazily
    p = f"mf_test_{typ}"
    users = g[f"{p}_users"].tolist()
    assert rec.plan.users == users
    ue, ie = t(g["user_emb"]), t(g["item_emb"])
    uid = ref_data.get_user_id_list(users)
    exact = _exact_scores_fn(ref_data, lambda j: (ue[uid[j]] @ ie.T).numpy(), users, typ, cold_object)
    O.check_topk_parity(g[f"{p}_scores"], g[f"{p}_dense_ids"], rec.scores.cpu().numpy(), rec.ids.cpu().numpy().astype(np.int64), exact)
    # the array path: no per-user Python work at all
    split = {"all": "overall_test", "cold": "cold_test", "warm": "warm_test"}[typ]
    plan = data.eval_plan(split, typ, cold_object, DEV)
    assert plan.users == users and torch.equal(plan.mask_col, rec.plan.mask_col) and torch.equal(plan.gt_col, rec.plan.gt_col)
    flags = cu(data.item_flags()) if cold_object == "item" else None
    s, i = FullRankScorer(20).topk([(tr.user_emb, tr.item_emb, None)], plan, flags)
    O.check_topk_parity(g[f"{p}_scores"], g[f"{p}_dense_ids"], s.cpu().numpy(), i.cpu().numpy().astype(np.int64), exact)
    if (g[f"{p}_scores"] > -1e8).all():
        got = device_metrics(i, plan.gt_rowptr, plan.gt_col, [10, 20], rounded=True)
        assert np.allclose(np.asarray(got), g[f"{p}_performance"], atol=1e-9)


def test_graph_and_sampler_from_array_builder():
    import coldrec_b200 as cr
    g, gg, gt = load_golden("eval_item"), load_golden("graph"), load_golden("train")
    data = cr.ArrayDataBuilder(*builder_args(g))
    G = data.graph(DEV)
    assert np.array_equal(G.rowptr.cpu().numpy(), gg["adj_indptr"]) and np.array_equal(G.col.cpu().numpy(), gg["adj_indices"])
    assert np.allclose(G.val.cpu().numpy(), gg["adj_data"], rtol=1e-6)
    u, i = cr.propagate(G, cu(gg["E0_user"]), cu(gg["E0_item"]), 3)
    assert (u.cpu() - t(gg["lgcn_L3_user"])).abs().max() <= 1e-5 * np.abs(gg["lgcn_L3_user"]).max()
    smp = data.sampler(DEV, seed=3)
    assert smp.n_pairs == len(gt["train_u"]) and smp.n_items == int(gt["n_item_table"])
    su, si, sj = (x.cpu().numpy().astype(np.int64) for x in smp.batch(0, 0, smp.n_pairs))
    O.check_sampler_epoch(su, si, sj, gt["train_u"], gt["train_i"], int(gt["n_item_table"]))
