"""Probe: the C3 (XING-shaped) towers, per layer — tcgen05 3xTF32 kernel vs the fp32 FFMA kernel vs torch (nn.Linear path of
the reference on the same GPU), with error against an fp64 evaluation.  Prints one JSON line per layer."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from coldrec_b200 import ops

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(4)
n_items, n_users, C, D = 20519, 106881, 2738, 64
content = (torch.rand(n_items, C, device=dev, generator=g) < 0.02).float() * torch.randn(n_items, C, device=dev, generator=g)
V = torch.randn(n_items, D, device=dev, generator=g) * 0.1
U = torch.randn(n_users, D, device=dev, generator=g) * 0.1
peak_tf32 = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["bf16_tflops"] / 2 if os.path.exists("MEASURED_PEAKS.json") else 795.0
hbm = 6545.9


def ms_of(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


cs = ops.split_tf32(content)
t_split = ms_of(lambda: ops.split_tf32(content))
print(json.dumps({"what": "split of the content table (once per table)", "ms": round(t_split, 3), "gbs": round(content.numel() * 12 / t_split / 1e6, 1)}), flush=True)
layers = [("DropoutNet item layer 1: [V | content] (20519 x 2802) -> 200, BN + tanh", V, content, 200),
          ("Heater gate + expert layer 1 stacked: content (20519 x 2738) -> 205, tanh", content, None, 205),
          ("DropoutNet item layer 2: 200 -> 100", torch.randn(n_items, 200, device=dev, generator=g), None, 100),
          ("DropoutNet user layer 1: 64 -> 200 over 106,881 users", U, None, 200),
          ("final 100 -> 64 over 106,881 users", torch.randn(n_users, 100, device=dev, generator=g), None, 64)]
for name, x1, x2, n_out in layers:
    k = x1.shape[1] + (0 if x2 is None else x2.shape[1])
    W = torch.randn(n_out, k, device=dev, generator=g) * (0.03 if k > 1000 else 0.08)
    b = torch.randn(n_out, device=dev, generator=g) * 0.05
    sc, sh = torch.rand(n_out, device=dev, generator=g) + 0.5, torch.randn(n_out, device=dev, generator=g) * 0.1
    s1 = cs if x1 is content else ops.split_tf32(x1)
    s2 = None if x2 is None else (cs if x2 is content else ops.split_tf32(x2))
    ws = ops.split_tf32(W)
    t_split_in = ms_of(lambda: ops.linear_act_tc(s1, ws, b, X2=s2, scale=sc, shift=sh, act="tanh", want_split=True))      # r02 first form: pre-split operands, split outputs
    r1 = ops.tma_rows(x1); r2 = None if x2 is None else ops.tma_rows(x2)
    t_tc = ms_of(lambda: ops.linear_act_tc(r1, ws, b, X2=r2, scale=sc, shift=sh, act="tanh"))                              # raw fp32 rows, split inside the kernel
    t_simt = ms_of(lambda: ops.linear_act(x1, W, b, X2=x2, scale=sc, shift=sh, act="tanh"), iters=3)
    xcat = x1 if x2 is None else torch.cat([x1, x2], 1)
    t_torch = ms_of(lambda: torch.tanh((torch.nn.functional.linear(xcat, W, b)) * sc + sh), iters=5)
    ref = torch.tanh((xcat.double() @ W.double().T + b.double()) * sc.double() + sh.double())
    y_tc = ops.linear_act_tc(r1, ws, b, X2=r2, scale=sc, shift=sh, act="tanh")[0]
    y_simt = ops.linear_act(x1, W, b, X2=x2, scale=sc, shift=sh, act="tanh")
    y_torch = torch.tanh((torch.nn.functional.linear(xcat, W, b)) * sc + sh)
    err = lambda y: float((y.double() - ref).abs().max() / ref.abs().max())
    flop = 2.0 * x1.shape[0] * k * n_out
    in_bytes = x1.shape[0] * k * 4 + x1.shape[0] * n_out * 4         # fp32 rows in, fp32 rows out
    print(json.dumps({"layer": name, "rows": x1.shape[0], "k": k, "n_out": n_out, "tc_ms": round(t_tc, 4), "tc_presplit_ms": round(t_split_in, 4), "simt_ms": round(t_simt, 4),
                      "torch_fp32_ms": round(t_torch, 4), "tc_tflops_algorithmic": round(flop / t_tc / 1e9, 1),
                      "tc_frac_of_tf32_peak_over_3": round(flop / t_tc / 1e9 / (peak_tf32 / 3), 3), "tc_gbs": round(in_bytes / t_tc / 1e6, 1),
                      "tc_frac_of_hbm": round(in_bytes / t_tc / 1e6 / hbm, 3), "err_tc": err(y_tc), "err_simt": err(y_simt), "err_torch_fp32": err(y_torch)}), flush=True)
