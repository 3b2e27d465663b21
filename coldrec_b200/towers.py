"""Content -> embedding generators (eval mode) on the fused tower kernel.

Each function takes the ``state_dict()`` of the corresponding reference module (tensors already on
the GPU, fp32) and reproduces its eval-mode forward with ``cr_linear_act_f32`` launches — Linear,
BatchNorm1d(running stats) and tanh fused in one kernel per layer, concat/gather/scatter folded into
the kernel's addressing:

  dropoutnet_encode  ``DeepCF.encode``            model/DropoutNet.py:192-213 (TanHBlock :222-236)
  heater_encode      ``Heater_encoder.encode``    model/Heater.py:187-223
  gar_generate       ``GAR_Learner.generator``    model/GAR.py:102-107, 130-131, cold-row overwrite :44-46
  aldi_tower         ``ALDITower.forward``        model/ALDI.py:191-208, 272-280
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import ops

State = Dict[str, torch.Tensor]


def _bn(state: State, prefix: str, eps: float):
    return ops.bn_fold(state.get(prefix + "weight"), state.get(prefix + "bias"), state[prefix + "running_mean"],
                       state[prefix + "running_var"], eps)


def _dropoutnet_tower(state: State, side: str, x1, x2):
    n_blocks = len({k.split(".")[1] for k in state if k.startswith(f"{side}_layers.")})
    h, h2 = x1, x2
    for l in range(n_blocks):
        p = f"{side}_layers.{l}."
        scale, shift = _bn(state, p + "bn.", 0.001)                       # BatchNorm1d(eps=0.001), DropoutNet.py:226-230
        h = ops.linear_act(h, state[p + "layer.weight"], state[p + "layer.bias"], X2=h2, scale=scale, shift=shift, act="tanh")
        h2 = None
    return ops.linear_act(h, state[f"{side}_emb.weight"], state[f"{side}_emb.bias"], X2=h2)


def dropoutnet_encode(state: State, Uin, Vin, Ucontent: Optional[torch.Tensor], Vcontent: Optional[torch.Tensor]):
    """(U_embedding, V_embedding) of ``DeepCF.encode``; content, when given, is the second K-segment
    of the first layer (the reference concatenates it after the CF embedding, :194-202)."""
    return _dropoutnet_tower(state, "u", Uin, Ucontent), _dropoutnet_tower(state, "v", Vin, Vcontent)


def heater_encode(state: State, Uin, Vin, Vcontent, n_expert: int, n_dropout: float):
    """Item-content branch of ``Heater_encoder.encode``.  The reference evaluates ``self.fc`` n_expert
    times on the same input (:191-193); the result is computed once and combined with the gate sum in
    ``cr_heater_blend_f32`` exactly as the bmm of :195 would."""
    gate = ops.linear_act(Vcontent, state["gate.linear.weight"], state["gate.linear.bias"], act="tanh")
    if gate.shape[1] != n_expert:
        raise ValueError(f"gate has {gate.shape[1]} outputs, n_expert={n_expert}")
    h = ops.linear_act(Vcontent, state["fc.linear1.weight"], state["fc.linear1.bias"], act="tanh")
    expert = ops.linear_act(h, state["fc.linear2.weight"], state["fc.linear2.bias"], act="tanh")
    keep = 1 - n_dropout                                                   # Vin_filter, :196
    v_last = ops.heater_blend(gate, expert, Vin.contiguous(), keep, 1 - keep)   # :195-198
    u_last = ops.linear_act(Uin, state["out_linear.weight"], state["out_linear.bias"], act="tanh")
    v_last = ops.linear_act(v_last, state["out_linear.weight"], state["out_linear.bias"], act="tanh")
    return (ops.linear_act(u_last, state["final_trans.weight"], state["final_trans.bias"]),
            ops.linear_act(v_last, state["final_trans.weight"], state["final_trans.bias"]))


def gar_generate(state: State, content, rows: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None):
    """tanh(L2(tanh(L1(content[rows])))).  With ``out`` given the result is scattered into
    ``out[rows]`` — the ``item_emb.data[cold_idx] = cold_item_gen_emb`` of GAR.py:44-46."""
    h = ops.linear_act(content, state["0.weight"], state["0.bias"], xrow=rows, act="tanh")
    return ops.linear_act(h, state["2.weight"], state["2.bias"], act="tanh", out=out, yrow=rows if out is not None else None)


def aldi_tower(state: State, x, rows: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None):
    """Linear -> BatchNorm1d (default eps 1e-5) -> tanh -> Linear, rows gathered / scattered like GAR."""
    scale, shift = _bn(state, "bn.", 1e-5)
    h = ops.linear_act(x, state["fc1.weight"], state["fc1.bias"], xrow=rows, scale=scale, shift=shift, act="tanh")
    return ops.linear_act(h, state["fc2.weight"], state["fc2.bias"], out=out, yrow=rows if out is not None else None)
