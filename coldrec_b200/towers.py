"""Content -> embedding generators (eval mode) on the fused tower kernels.

Each function takes the ``state_dict()`` of the corresponding reference module (tensors already on
the GPU, fp32) and reproduces its eval-mode forward — Linear, BatchNorm1d(running stats) and tanh fused
in one kernel per layer, concat / gather / scatter folded into the kernel's addressing:

  dropoutnet_encode  ``DeepCF.encode``            model/DropoutNet.py:192-213 (TanHBlock :222-236)
  heater_encode      ``Heater_encoder.encode``    model/Heater.py:187-223
  gar_generate       ``GAR_Learner.generator``    model/GAR.py:102-107, 130-131, cold-row overwrite :44-46
  aldi_tower         ``ALDITower.forward``        model/ALDI.py:191-208, 272-280

Layers run on the tensor cores (``cr_linear_act_tc_f32``: tcgen05 kind::tf32 with every operand split hi + lo and
hi.hi + lo.hi + hi.lo accumulated in fp32 — fp32 accuracy, the 1e-5 norm-wise parity bar holds); a chain hands its
activations on already split, so only the first input of a tower is split by a pass of its own.  A CONSTANT input — the item
content matrix, 225 MB at XING shape — should be split once with ``prepare(content)`` and the result passed wherever the
functions below take content.  ``CR_TOWERS_SIMT=1`` routes everything to the fp32 FFMA kernel (``cr_linear_act_f32``) instead.
"""
from __future__ import annotations

import os
from typing import Dict, Optional, Union

import torch

from . import ops
from .ops import SplitTable

State = Dict[str, torch.Tensor]
Table = Union[torch.Tensor, SplitTable]


def prepare(table: torch.Tensor) -> SplitTable:
    """Split a constant fp32 table (item / user content) once for the tensor-core towers."""
    return ops.split_tf32(table)


def _use_tc(n_out: int) -> bool:
    return n_out <= 256 and not os.environ.get("CR_TOWERS_SIMT")


def _plain(t: Table) -> torch.Tensor:
    if isinstance(t, SplitTable):
        return (t.hi + t.lo)[:, :t.width]
    return t


def _split(t: Table, rows: Optional[torch.Tensor] = None) -> SplitTable:
    """Operand form of a table; with ``rows`` only those rows (the ``content[cold_idx]`` gather of GAR.py:44-46)."""
    if isinstance(t, SplitTable):
        if rows is None:
            return t
        return SplitTable(ops.gather_rows(t.hi, rows), ops.gather_rows(t.lo, rows), t.width)
    if rows is not None:
        t = ops.gather_rows(t.contiguous(), rows)
    return ops.split_tf32(t)


def _layer(x1: Table, weight, bias, *, x2: Optional[Table] = None, xrow=None, scale=None, shift=None, act=None, out=None, yrow=None,
           want_split: bool = False):
    """One fused layer.  Returns the fp32 output, or — with ``want_split`` — its SplitTable for the next layer."""
    if _use_tc(weight.shape[0]):
        y, sp = ops.linear_act_tc(_split(x1, xrow), ops.split_tf32(weight), bias, X2=None if x2 is None else _split(x2, xrow), scale=scale,
                                  shift=shift, act=act, out=out, yrow=yrow, want_split=want_split, want_plain=not want_split)
        return sp if want_split else y
    return ops.linear_act(_plain(x1), weight, bias, X2=None if x2 is None else _plain(x2), xrow=xrow, scale=scale, shift=shift, act=act,
                          out=out, yrow=yrow)


def _bn(state: State, prefix: str, eps: float):
    return ops.bn_fold(state.get(prefix + "weight"), state.get(prefix + "bias"), state[prefix + "running_mean"],
                       state[prefix + "running_var"], eps)


def _dropoutnet_tower(state: State, side: str, x1, x2):
    n_blocks = len({k.split(".")[1] for k in state if k.startswith(f"{side}_layers.")})
    h, h2 = x1, x2
    for l in range(n_blocks):
        p = f"{side}_layers.{l}."
        scale, shift = _bn(state, p + "bn.", 0.001)                       # BatchNorm1d(eps=0.001), DropoutNet.py:226-230
        h = _layer(h, state[p + "layer.weight"], state[p + "layer.bias"], x2=h2, scale=scale, shift=shift, act="tanh", want_split=True)
        h2 = None
    return _layer(h, state[f"{side}_emb.weight"], state[f"{side}_emb.bias"], x2=h2)


def dropoutnet_encode(state: State, Uin, Vin, Ucontent: Optional[Table], Vcontent: Optional[Table]):
    """(U_embedding, V_embedding) of ``DeepCF.encode``; content, when given, is the second K-segment
    of the first layer (the reference concatenates it after the CF embedding, :194-202)."""
    return _dropoutnet_tower(state, "u", Uin, Ucontent), _dropoutnet_tower(state, "v", Vin, Vcontent)


def heater_encode(state: State, Uin, Vin, Vcontent: Table, n_expert: int, n_dropout: float):
    """Item-content branch of ``Heater_encoder.encode``.  The reference evaluates ``self.fc`` n_expert
    times on the same input (:191-193); the result is computed once and combined with the gate sum in
    ``cr_heater_blend_f32`` exactly as the bmm of :195 would.  The gate and the first expert layer read the same content and
    both end in tanh: they run as ONE layer over the stacked weights (one pass over the content instead of two)."""
    n_gate = state["gate.linear.weight"].shape[0]
    if n_gate != n_expert:
        raise ValueError(f"gate has {n_gate} outputs, n_expert={n_expert}")
    n_h = state["fc.linear1.weight"].shape[0]
    if _use_tc(n_h + n_gate):
        w = torch.cat([state["fc.linear1.weight"], state["gate.linear.weight"]], 0)
        b = torch.cat([state["fc.linear1.bias"], state["gate.linear.bias"]], 0)
        hg = _layer(Vcontent, w, b, act="tanh")                                         # [tanh(fc1) | tanh(gate)]
        h, gate = hg[:, :n_h], hg[:, n_h:].contiguous()
    else:
        gate = _layer(Vcontent, state["gate.linear.weight"], state["gate.linear.bias"], act="tanh")
        h = _layer(Vcontent, state["fc.linear1.weight"], state["fc.linear1.bias"], act="tanh")
    expert = _layer(h, state["fc.linear2.weight"], state["fc.linear2.bias"], act="tanh")
    keep = 1 - n_dropout                                                   # Vin_filter, :196
    v_last = ops.heater_blend(gate, expert, Vin.contiguous(), keep, 1 - keep)   # :195-198
    u_last = _layer(Uin, state["out_linear.weight"], state["out_linear.bias"], act="tanh", want_split=True)
    v_last = _layer(v_last, state["out_linear.weight"], state["out_linear.bias"], act="tanh", want_split=True)
    return (_layer(u_last, state["final_trans.weight"], state["final_trans.bias"]),
            _layer(v_last, state["final_trans.weight"], state["final_trans.bias"]))


def gar_generate(state: State, content: Table, rows: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None):
    """tanh(L2(tanh(L1(content[rows])))).  With ``out`` given the result is scattered into
    ``out[rows]`` — the ``item_emb.data[cold_idx] = cold_item_gen_emb`` of GAR.py:44-46."""
    h = _layer(content, state["0.weight"], state["0.bias"], xrow=rows, act="tanh", want_split=True)
    return _layer(h, state["2.weight"], state["2.bias"], act="tanh", out=out, yrow=rows if out is not None else None)


def aldi_tower(state: State, x: Table, rows: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None):
    """Linear -> BatchNorm1d (default eps 1e-5) -> tanh -> Linear, rows gathered / scattered like GAR."""
    scale, shift = _bn(state, "bn.", 1e-5)
    h = _layer(x, state["fc1.weight"], state["fc1.bias"], xrow=rows, scale=scale, shift=shift, act="tanh", want_split=True)
    return _layer(h, state["fc2.weight"], state["fc2.bias"], out=out, yrow=rows if out is not None else None)
