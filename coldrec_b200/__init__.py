"""coldrec_b200 — B200-native embedding-generation-and-scoring path of ColdRec (sm_100a only).

Host code mirrors the reference's interfaces (``BaseColdStartTrainer``, ``ranking_evaluation``, the
propagation loop of the graph encoders, the generator towers) and calls hand-written CUDA kernels
through the C ABI declared in ``include/coldrec_b200.h``.  No CPU fallback.
"""
from . import ops
from .evaluator import RecList, ranking_evaluation
from .graph import (CsrGraph, PropagationBuffers, bipartite_norm_csr, propagate, propagate_frozen_cold, propagate_ngcf,
                    propagate_table)
from .scoring import FLAG_COLD, FLAG_WARM, EvalPlan, FullRankScorer, item_flags_from
from .trainer import AldiScoreTables, BaseColdStartTrainer, FusedEvalMixin, TwoProductScoreTables
from . import towers
from .training import BprTrainStep, PairwiseSampler
from .databuilder import ArrayDataBuilder

__all__ = ["ops", "towers", "RecList", "ranking_evaluation", "CsrGraph", "bipartite_norm_csr", "propagate", "propagate_ngcf", "propagate_frozen_cold",
           "FLAG_COLD", "FLAG_WARM", "EvalPlan", "FullRankScorer", "item_flags_from", "AldiScoreTables",
           "BaseColdStartTrainer", "FusedEvalMixin", "TwoProductScoreTables", "PropagationBuffers", "propagate_table",
           "BprTrainStep", "PairwiseSampler", "ArrayDataBuilder"]
