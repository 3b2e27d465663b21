"""The drop-in claim, proven on the reference's OWN classes (baseline/_ref = the unmodified ColdRec tree).

  * CPU: the early-stopping state machine / printed progress block of ``FusedEvalMixin.fast_evaluation`` equals
    ``model/BaseRecommender.py:268-351`` for scripted validation results (improving, equal, worse, non-finite).
  * GPU: ``class MF(FusedEvalMixin, reference.MF)`` and ``LightGCN`` (encoder forward swapped for ``CsrGraph`` +
    ``propagate`` at eval time) are trained for 2 epochs by the reference's own ``train()`` and
    taken through ``run()``; the per-epoch validation lines and the results of all three settings must equal what the
    reference's own ``_evaluate`` + ``util.evaluator.ranking_evaluation`` give for the same tables; a replay of fixed
    tables compares the early-stop counters with the reference's ``fast_evaluation`` / ``run()`` executed on the CPU.
"""
import argparse
import contextlib
import io
import types

import numpy as np
import pytest
import torch

from tests import refimport
from tests.helpers import builder_args, load_golden

pytestmark = pytest.mark.skipif(not refimport.available(), reason="baseline/_ref missing (python baseline/install_ref.py)")


def _args(model, epochs, early_stop, cold_object="item", bs=256, layers=2):
    return argparse.Namespace(topN="10,20", model=model, dataset="syn", emb_size=64, epochs=epochs, bs=bs, lr=5e-3, reg=1e-4,
                              early_stop=early_stop, eval_every=1, cold_object=cold_object, save_emb=False, layers=layers, seed=7)


class _Cfg:
    def __init__(self, args, data, device):
        self.args, self.data, self.device = args, data, torch.device(device)


def _quiet(fn, *a, **kw):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        out = fn(*a, **kw)
    return out, buf.getvalue()


# ---------------------------------------------------------------------------------------------- CPU: state machine
def _measure(ndcg, hr=0.1):
    return ['Top 20\n', f'Hit Ratio:{hr}\n', 'Precision:0.01\n', 'Recall:0.2\n', f'NDCG:{ndcg}\n']


SCRIPTS = {
    "improve_then_stall": [0.10, 0.12, 0.12, 0.11, 0.13, 0.05, 0.05, 0.05],
    "nan_first": [float("nan"), float("nan"), 0.2, float("inf"), 0.1, 0.3],
    "nan_later": [0.2, float("nan"), 0.2, 0.25, float("nan"), float("nan"), float("nan")],
}


@pytest.mark.parametrize("script", sorted(SCRIPTS))
@pytest.mark.parametrize("early_stop", [0, 2])
def test_fast_evaluation_state_machine_equals_reference(script, early_stop, monkeypatch):
    refimport.import_reference()
    import model.BaseRecommender as RB
    from coldrec_b200.trainer import BaseColdStartTrainer as Mine

    data = types.SimpleNamespace(overall_valid_set={1: {2: 1.0}}, warm_valid_set={}, cold_valid_set={})
    seq = SCRIPTS[script]

    def make(base, patch):
        class T(base):
            saves = 0
            def train(self): pass
            def predict(self, u): pass
            def batch_predict(self, users): pass
            def save(self): type(self).saves += 1
            def valid(self, valid_type='all'): return {1: [(2, 1.0)]}
        t = T(_Cfg(_args("MF", len(seq), early_stop), data, "cpu"))
        it = iter(seq)
        patch(t, lambda gt, rec, N: (_measure(next(it)), [[0.0] * 4]))
        return t

    ref = make(RB.BaseColdStartTrainer, lambda t, f: monkeypatch.setattr(RB, "ranking_evaluation", f))
    ref_log = []
    for e in range(len(seq)):
        lines, out = _quiet(ref.fast_evaluation, e)
        ref_log.append((lines, out, list(ref.bestPerformance), getattr(ref, "early_stop_patience", None), type(ref).saves))
    mine = make(Mine, lambda t, f: setattr(t, "_ranking_evaluation", f))
    for e in range(len(seq)):
        lines, out = _quiet(mine.fast_evaluation, e)
        want = ref_log[e]
        assert lines == want[0] and out == want[1], f"epoch {e}: printed block differs"
        assert list(mine.bestPerformance) == want[2] or (np.isnan(seq[e]) and str(mine.bestPerformance) == str(want[2]))
        assert getattr(mine, "early_stop_patience", None) == want[3] and type(mine).saves == want[4]
    with pytest.raises(ValueError, match="Invalid evaluation type!"):
        mine.fast_evaluation(0, valid_type="bogus")


# ---------------------------------------------------------------------------------------------- GPU: grafted classes
def _reference_data(name="eval_item"):
    refimport.import_reference()
    from util.databuilder import ColdStartDataBuilder
    g = load_golden(name)
    return ColdStartDataBuilder(*builder_args(g)), str(g["cold_object"])


def _reference_twin(trainer, ref_cls):
    """An instance of the UNTOUCHED reference class sharing the grafted trainer's state (tables, data, args): every method
    it runs — _get_eval_cache, _evaluate, batch_predict — resolves in the reference's own MRO, none in the mixin's."""
    twin = object.__new__(ref_cls)
    twin.__dict__.update({k: v for k, v in trainer.__dict__.items() if k != "_fused"})
    twin._eval_cache = {}
    assert "FusedEvalMixin" not in [c.__name__ for c in type(twin).__mro__]
    return twin


def _reference_eval(trainer, ref_cls, kind, split, rec_only=False):
    """The reference's own _evaluate (its torch path on the trainer's device) + its own ranking_evaluation."""
    from util.evaluator import ranking_evaluation
    twin = _reference_twin(trainer, ref_cls)
    gt = getattr(twin.data, f"{ {'all': 'overall'}.get(kind, kind) }_{split}_set")
    rec = twin._evaluate(gt, kind)
    return rec if rec_only else ranking_evaluation(gt, rec, twin.topN if split == "test" else [twin.max_N])


def _run_grafted(cls_name, make_cls, epochs=2):
    from coldrec_b200 import FusedEvalMixin
    data, cold_object = _reference_data()
    Fused = make_cls(FusedEvalMixin)
    torch.manual_seed(7); np.random.seed(7)
    import random
    random.seed(7)
    t = Fused(_Cfg(_args(cls_name, epochs, 0, cold_object), data, "cuda:0"))
    _, out = _quiet(t.run)
    return t, out


@pytest.mark.gpu
def test_mixin_on_reference_mf_run_equals_reference_evaluation():
    refimport.import_reference()
    from model.MF import MF as RefMF

    def make(Mixin):
        class MF(Mixin, RefMF):
            epoch_tables, epoch_lines = [], []
            def fast_evaluation(self, epoch, valid_type='all'):
                type(self).epoch_tables.append((self.user_emb.detach().clone(), self.item_emb.detach().clone()))
                lines = super().fast_evaluation(epoch, valid_type)
                type(self).epoch_lines.append(lines)
                return lines
        return MF
    t, out = _run_grafted("MF", make)
    assert type(t).__mro__[1].__name__ == "FusedEvalMixin" and t.epochs_ran == 2
    assert "Testing under [cold] setting..." in out and "[warm setting] The result of MF:" in out
    # final results of the three settings == the reference's own evaluation of the same (best) tables
    for kind, attr in (("all", "overall_test_results"), ("cold", "cold_test_results"), ("warm", "warm_test_results")):
        measure, perf = _reference_eval(t, RefMF, kind, "test")
        assert getattr(t, attr) == perf, f"{kind}: {getattr(t, attr)} vs reference {perf}"
        assert f"[{kind} setting] The result of MF:\n{''.join(measure)}" in out
    # per-epoch validation lines == the reference's, on the tables of that epoch
    for (ue, ie), lines in zip(type(t).epoch_tables, type(t).epoch_lines):
        t.user_emb, t.item_emb = ue, ie
        measure, _ = _reference_eval(t, RefMF, "all", "valid")
        assert lines == [m.strip() for m in measure[1:]]
    # the lazily built rec list is the reference's dict format
    t.user_emb, t.item_emb = t.best_user_emb, t.best_item_emb
    rec = t.test("all")
    ref_rec = _reference_eval(t, RefMF, "all", "test", rec_only=True)
    assert list(rec.keys()) == list(ref_rec.keys())
    u0 = next(iter(ref_rec))
    assert [i for i, _ in rec[u0]] == [i for i, _ in ref_rec[u0]]
    assert np.allclose([s for _, s in rec[u0]], [s for _, s in ref_rec[u0]], atol=1e-5)


@pytest.mark.gpu
def test_mixin_on_reference_lightgcn_with_csr_graph_swapped_in():
    refimport.import_reference()
    import importlib
    RL = importlib.import_module("model.LightGCN")            # (`model.LightGCN` the attribute is the trainer class: model/__init__.py)
    from coldrec_b200 import CsrGraph, propagate

    class FusedEncoder(RL.LGCN_Encoder):                       # the eval-time forward on the fused SpMM
        def __init__(self, *a):
            super().__init__(*a)
            self.graph = CsrGraph.from_scipy(self.norm_adj, self.device)
        def forward(self):
            if torch.is_grad_enabled():                        # training keeps the autograd path (torch.sparse.mm)
                return super().forward()
            return propagate(self.graph, self.embedding_dict['user_emb'].detach(), self.embedding_dict['item_emb'].detach(), self.layers)

    def make(Mixin):
        class LightGCN(Mixin, RL.LightGCN):
            def __init__(self, config):
                super().__init__(config)
                torch.manual_seed(11)
                self.model = FusedEncoder(self.data, self.emb_size, self.n_layers, self.device)
        return LightGCN
    t, out = _run_grafted("LightGCN", make)
    with torch.no_grad():                                      # the reference encoder's own forward on the trained parameters
        ref_u, ref_i = RL.LGCN_Encoder.forward(t.model)
    for got, ref in ((t.best_user_emb, ref_u), (t.best_item_emb, ref_i)):
        assert (got - ref).abs().max().item() <= 1e-5 * ref.abs().max().item()
    t.user_emb, t.item_emb = ref_u, ref_i
    for kind, attr in (("all", "overall_test_results"), ("cold", "cold_test_results"), ("warm", "warm_test_results")):
        _, perf = _reference_eval(t, RL.LightGCN, kind, "test")
        assert getattr(t, attr) == perf, f"{kind}: {getattr(t, attr)} vs reference {perf}"


@pytest.mark.gpu
@pytest.mark.parametrize("golden", ["eval_item", "eval_user"])
def test_replayed_tables_early_stop_counters_equal_reference_cpu_run(golden):
    """Fixed per-epoch tables through the grafted class on the GPU and through the untouched reference class on the CPU:
    printed validation lines, best epoch, patience and the final three-setting results must be identical."""
    refimport.import_reference()
    from model.MF import MF as RefMF
    from coldrec_b200 import FusedEvalMixin
    data, cold_object = _reference_data(golden)
    g = torch.Generator().manual_seed(3)
    A = (torch.randn(data.user_num, 64, generator=g) * 0.1, torch.randn(data.item_num, 64, generator=g) * 0.1)
    gl = load_golden(golden)
    B = (torch.from_numpy(gl["user_emb"]), torch.from_numpy(gl["item_emb"]))
    C = (B[0] * 1.0, B[1] + 0.05 * A[1])
    tables = [A, B, A, C, A, A, A, B]                      # improves, falls back, may improve, then stalls until patience runs out

    def make(*bases):
        class Replay(*bases):
            def train(self):
                for epoch, (ue, ie) in enumerate(tables):
                    self.user_emb, self.item_emb = ue.to(self.device), ie.to(self.device)
                    self.fast_evaluation(epoch)
                    if self.early_stop_flag and self.early_stop_patience <= 0:
                        break
                self.epochs_ran = epoch + 1
                self.user_emb, self.item_emb = self.best
            def save(self):
                self.best = (self.user_emb.clone(), self.item_emb.clone())
        return Replay
    res = {}
    for name, cls, dev in (("ref", make(RefMF), "cpu"), ("fused", make(FusedEvalMixin, RefMF), "cuda:0")):
        t = cls(_Cfg(_args("MF", len(tables), 3, cold_object, bs=64), data, dev))
        _, out = _quiet(t.run)
        res[name] = (out, t.bestPerformance, t.early_stop_patience, t.epochs_ran, t.overall_test_results, t.cold_test_results,
                     t.warm_test_results)
    assert res["fused"][1:] == res["ref"][1:]
    assert res["fused"][0] == res["ref"][0], "printed output of run() differs from the reference's"
    assert res["ref"][3] < len(tables), "the scenario must actually stop early"
