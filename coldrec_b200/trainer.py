"""The reference's trainer API (model/BaseRecommender.py) with the evaluation hot path on the GPU.

``BaseColdStartTrainer`` keeps the constructor, attributes and method names of the reference class
(``train`` / ``predict`` / ``batch_predict`` / ``save`` abstract; ``_evaluate`` / ``valid`` / ``test`` /
``full_evaluation`` / ``fast_evaluation`` / ``run`` concrete) so the 26 ``model/*.py`` trainers can
subclass it unchanged.  Differences, all inside the hot path:

  * ``_evaluate`` (reference :153-188) does not call ``batch_predict``; it asks the model for its score
    tables (``get_score_tables``, by default ``[(get_user_emb(), get_item_emb(), None)]``) and runs the
    fused scorer, returning a lazily materialised ``RecList``.
  * ``ranking_evaluation`` is the device reduction of ``coldrec_b200.evaluator``.
  * ``get_user_emb`` / ``get_item_emb`` are added as accessors over ``self.user_emb`` / ``self.item_emb``
    (in the reference only USIMCore has them, model/USIM.py:602,628).

``FusedEvalMixin`` carries the same overrides for grafting onto the reference's own base class:
``class MF(FusedEvalMixin, reference.BaseColdStartTrainer)`` — see INTEGRATION.md.
"""
from __future__ import annotations

import math
import time
from abc import ABC, abstractmethod
from typing import Dict, List, Tuple

import torch

from . import ops
from .evaluator import RecList, ranking_evaluation
from .scoring import FLAG_COLD, FLAG_WARM, GROUP_UNFLAGGED, EvalPlan, FullRankScorer, item_flags_from


class FusedEvalMixin:
    """Overrides of the evaluation path; expects the attributes set by BaseColdStartTrainer.__init__."""

    score_precision = ops.SCORE_TF32_CHECKED

    # ---- accessors named by the north star -----------------------------------------------------
    def get_user_emb(self) -> torch.Tensor:
        return self.user_emb

    def get_item_emb(self) -> torch.Tensor:
        return self.item_emb

    def get_score_tables(self):
        """[(user_table, item_table, item_group)] — see FullRankScorer.  Models whose ``batch_predict`` is
        not a single inner product override this (ALDI, VBPR/AMR below)."""
        return [(self.get_user_emb(), self.get_item_emb(), None)]

    # ---- the hot path ---------------------------------------------------------------------------
    def _fused_state(self):
        st = self.__dict__.get("_fused")
        if st is None:
            dev = torch.device(self.device)
            if dev.type != "cuda":
                raise RuntimeError("coldrec_b200 evaluates on a B200 only: config.device must be a CUDA device "
                                   "(there is no CPU fallback)")
            st = {"scorer": FullRankScorer(self.max_N, self.score_precision), "plans": {},
                  "flags": item_flags_from(self.data, dev) if self.args.cold_object == 'item' else None, "device": dev}
            self.__dict__["_fused"] = st
        return st

    def _get_eval_cache(self, data_set: Dict, data_type: str) -> EvalPlan:
        """Reference :109-151, memoised on the same key; one CSR + flags instead of per-user tensors."""
        st = self._fused_state()
        key = (id(data_set), data_type, str(self.device), self.args.cold_object)
        plan = st["plans"].get(key)
        if plan is None:
            # an ArrayDataBuilder knows the split this dict is a view of: its plan comes from arrays, no per-user Python work
            split = self.data.split_of(data_set) if hasattr(self.data, "split_of") else None
            if split is not None:
                plan = self.data.eval_plan(split, data_type, self.args.cold_object, st["device"])
            else:
                plan = EvalPlan.from_data(self.data, data_set, data_type, self.args.cold_object, st["device"])
            st["plans"][key] = plan
        return plan

    def _evaluate(self, data_set: Dict, data_type: str = 'all') -> RecList:
        """Reference :153-188: full ranking of every eval user -> {user: [(item, score) x max_N]}."""
        st = self._fused_state()
        plan = self._get_eval_cache(data_set, data_type)
        tables = []
        for ut, it, grp in self.get_score_tables():
            ut = ut.detach().to(device=st["device"], dtype=torch.float32).contiguous()
            it = it.detach().to(device=st["device"], dtype=torch.float32).contiguous()
            tables.append((ut, it, grp))
        scores, ids = st["scorer"].topk(tables, plan, st["flags"])
        rec = RecList(plan, scores, ids, self.data.id2item)
        rec._gt_id = id(data_set)
        return rec

    def _ranking_evaluation(self, gt, rec_list, topN):
        return ranking_evaluation(gt, rec_list, topN, item_map=self.data.item, device=self._fused_state()["device"])

    # ---- the consumers of the hot path (reference :230-351).  They are overridden here, not only in the base class below,
    # because the reference's own versions call the module-level ``util.evaluator.ranking_evaluation`` (:7), which would
    # materialise every RecList into Python dicts and run the 24 us/user metric loops on the host again. -------------------
    _SPLIT_NAMES = {'warm': 'warm', 'cold': 'cold', 'all': 'overall'}

    def _split(self, prefix: str, kind: str, what: str):
        if kind not in self._SPLIT_NAMES:
            raise ValueError(f'Invalid {what} type!')
        return getattr(self.data, f'{self._SPLIT_NAMES[kind]}_{prefix}_set')

    def full_evaluation(self, rec_list, test_type: str = 'warm') -> None:
        """Reference :230-254: metrics at every cut-off of ``topN`` on the test split, kept in ``<setting>_test_results``."""
        test_set = self._split('test', test_type, 'evaluation')
        self.result, performance = self._ranking_evaluation(test_set, rec_list, self.topN)
        setattr(self, self._SPLIT_NAMES[test_type] + '_test_results', performance)
        print('*' * 80)
        print(f'[{test_type} setting] The result of %s:\n%s' % (self.model_name, ''.join(self.result)))

    @staticmethod
    def _metrics_dict_from_measure(measure: List[str]) -> Dict[str, float]:
        """'NDCG:0.123\n' lines -> {'NDCG': 0.123} (reference :256-262; the format ``ranking_evaluation`` must keep)."""
        return {k: float(v) for k, v in (m.strip().split(':') for m in measure[1:])}

    @staticmethod
    def _metrics_all_finite(performance: Dict[str, float]) -> bool:
        return all(math.isfinite(v) for v in performance.values())

    def _track_best(self, epoch: int, performance: Dict[str, float]) -> None:
        """Early-stopping state machine of reference :291-327.  A validation counts as an improvement iff its metrics are
        finite and NDCG@max(topN) is strictly above the best so far (any finite validation, when there is no best yet);
        an improvement saves the model and — only when a best already existed — refills the patience; anything else costs
        one unit of patience."""
        had_best = len(self.bestPerformance) > 0
        finite = self._metrics_all_finite(performance)
        if finite and (not had_best or performance['NDCG'] > self.bestPerformance[1]['NDCG']):
            self.bestPerformance = [epoch + 1, performance]
            self.save()
            if had_best and self.early_stop_flag:
                self.early_stop_patience = self.max_early_stop_patience
            return
        if self.early_stop_flag:
            self.early_stop_patience -= 1
        if not finite and had_best:
            print('Warning: validation metrics are non-finite; early-stop patience decreased, best checkpoint unchanged.')
        elif not finite and self.early_stop_flag:
            print('Warning: first validation has non-finite metrics; best checkpoint not initialized yet.')

    def fast_evaluation(self, epoch: int, valid_type: str = 'all') -> List[str]:
        """Reference :268-351: validate at max(topN), update the early-stopping state, print the progress block."""
        valid_set = self._split('valid', valid_type, 'evaluation')
        print(f'Evaluating the model under the {valid_type} setting...')
        measure, _ = self._ranking_evaluation(valid_set, self.valid(valid_type), [self.max_N])
        self._track_best(epoch, self._metrics_dict_from_measure(measure))
        lines = [m.strip() for m in measure[1:]]
        rule = '-' * 120
        print(rule)
        print('Performance ' + ' (Top-' + str(self.max_N) + ' Recommendation)')
        print('*Current Performance*')
        print('Epoch:', str(epoch + 1) + ',', '  |  '.join(lines))
        if self.bestPerformance:
            best_epoch, best = self.bestPerformance
            print(f'*Best {valid_type} Performance* ')
            print('Epoch:', str(best_epoch) + ',', '  |  '.join(f'{k}:{best[k]}' for k in ('Hit Ratio', 'Precision', 'Recall', 'NDCG')))
        else:
            print(f'*Best {valid_type} Performance* not initialized (waiting for finite validation).')
        if self.early_stop_flag:
            print(f"Stopping early at epoch {epoch + 1}." if self.early_stop_patience <= 0
                  else f"Early stopping patience left: {self.early_stop_patience}.")
        print(rule)
        return lines


class BaseColdStartTrainer(FusedEvalMixin, ABC):
    """Mirror of model/BaseRecommender.py:13-370 (same constructor contract: ``config.args``,
    ``config.data``, ``config.device``) for trainers that do not inherit the reference's class."""

    def __init__(self, config):
        self.config, self.args, self.data, self.device = config, config.args, config.data, config.device
        a = self.args
        self.model_name, self.dataset_name, self.emb_size = a.model, a.dataset, a.emb_size
        self.maxEpoch, self.batch_size, self.lr, self.reg = a.epochs, a.bs, a.lr, a.reg
        self.topN = [int(n) for n in a.topN.split(',')]
        self.max_N = max(self.topN)
        self.bestPerformance, self.result, self.epochs_ran = [], [], 0
        self.early_stop_flag = a.early_stop != 0
        if self.early_stop_flag:
            self.early_stop_patience = self.max_early_stop_patience = a.early_stop
        self.eval_every = max(1, int(getattr(a, 'eval_every', 1)))

    def print_basic_info(self):
        print('*' * 80)
        for k, v in (('Model: ', self.model_name), ('Dataset: ', self.dataset_name), ('Embedding Dimension:', self.emb_size),
                     ('Maximum Epoch:', self.maxEpoch), ('Learning Rate:', self.lr), ('Batch Size:', self.batch_size)):
            print(k, v)
        print('*' * 80)

    def timer(self, start=True):
        setattr(self, 'train_start_time' if start else 'train_end_time', time.time())

    @abstractmethod
    def train(self) -> None: ...

    @abstractmethod
    def save(self) -> None: ...

    def predict(self, u):
        """Scores of one user over all items (reference abstract :74; MF.py:52-56 given here as default)."""
        with torch.no_grad():
            uid = self.data.get_user_id(u)
            return torch.matmul(self.get_user_emb()[uid], self.get_item_emb().transpose(0, 1)).cpu().numpy()

    def batch_predict(self, users):
        """Dense (len(users), item_num) scores (reference abstract :87).  Kept for API compatibility with
        callers outside evaluation; ``_evaluate`` never materialises this matrix."""
        with torch.no_grad():
            uids = torch.as_tensor(self.data.get_user_id_list(users), device=self.get_user_emb().device)
            return torch.matmul(self.get_user_emb()[uids], self.get_item_emb().transpose(0, 1))

    def valid(self, valid_type: str = 'all'):
        return self._evaluate(self._split('valid', valid_type, 'valid'), valid_type)

    def test(self, test_type: str = 'all'):
        return self._evaluate(self._split('test', test_type, 'test'), test_type)

    def run(self) -> None:
        """Reference :353-370: train, then test + evaluate under all / cold / warm."""
        self.print_basic_info()
        print('Training Model...')
        self.train()
        if getattr(self, 'epochs_ran', 0) == 0 and self.maxEpoch > 0:
            self.epochs_ran = self.maxEpoch
        for setting in ('all', 'cold', 'warm'):
            print('*' * 80)
            print(f'Testing under [{setting}] setting...')
            rec_list = self.test(test_type=setting)
            print(f'Evaluating under [{setting}] setting...')
            self.full_evaluation(rec_list, test_type=setting)


class AldiScoreTables:
    """``get_score_tables`` for ALDI (model/ALDI.py:149-160): warm items are scored with
    ``warm_user_emb``, cold items with ``cold_user_emb``; rows of the item table in neither list keep
    the zero the reference initialises the score matrix with."""

    def get_score_tables(self):
        tables = [(self.warm_user_emb, self.item_emb, FLAG_WARM), (self.cold_user_emb, self.item_emb, FLAG_COLD)]
        flags = self._fused_state()["flags"]
        if bool(((flags & (FLAG_WARM | FLAG_COLD)) == 0).any()):
            tables.append((torch.zeros_like(self.warm_user_emb), self.item_emb, GROUP_UNFLAGGED))
        return tables


class TwoProductScoreTables:
    """``get_score_tables`` for VBPR / AMR (model/VBPR.py:68-75, model/AMR.py:68-76):
    P.Q^T + P2.Q2^T == [P|P2].[Q|Q2]^T, one sweep at d = d1 + d2."""

    def get_score_tables(self):
        return [(torch.cat([self.user_emb_main, self.user_emb_aux], 1), torch.cat([self.item_emb_main, self.item_emb_aux], 1), None)]
