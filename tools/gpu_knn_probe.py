"""Probe: content kNN at XING shape (model/KNN.py:63-77: 4,104 cold items x 16,415 warm items x 2,738-d content, k = 5..20) —
tensor-core path (3xTF32 GEMM blocks + row top-k) vs the exact fp32 FFMA sweep vs torch.matmul + topk on the same GPU."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from coldrec_b200 import knn

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(9)


def ms_of(fn, iters=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


for name, n_q, n_v, d, k in (("XING", 4104, 16415, 2738, 10), ("CiteULike", 3396, 13584, 300, 10)):
    Q = (torch.rand(n_q, d, device=dev, generator=g) < 0.05).float() * torch.randn(n_q, d, device=dev, generator=g)
    V = (torch.rand(n_v, d, device=dev, generator=g) < 0.05).float() * torch.randn(n_v, d, device=dev, generator=g)
    t_tc = ms_of(lambda: knn.knn_inner_product(Q, V, k, dev))
    os.environ["CR_KNN_SIMT"] = "1"
    t_simt = ms_of(lambda: knn.knn_inner_product(Q, V, k, dev), iters=2)
    s1, i1 = knn.knn_inner_product(Q, V, k, dev)
    del os.environ["CR_KNN_SIMT"]
    s0, i0 = knn.knn_inner_product(Q, V, k, dev)
    t_torch = ms_of(lambda: torch.topk(Q @ V.T, k, dim=1))
    ts, ti = torch.topk(Q.double() @ V.double().T, k, dim=1)
    flop = 2.0 * n_q * n_v * d
    print(json.dumps({"shape": name, "n_query": n_q, "n_value": n_v, "d": d, "k": k, "tc_ms": round(t_tc, 3), "exact_simt_ms": round(t_simt, 3),
                      "torch_matmul_topk_ms": round(t_torch, 3), "tc_tflops_algorithmic": round(flop / t_tc / 1e9, 1),
                      "ids_equal_fp64": float((i0.long() == ti).float().mean()), "ids_equal_simt": float((i0 == i1).float().mean()),
                      "max_score_err_vs_fp64": float((s0.double() - ts).abs().max() / ts.abs().max())}), flush=True)
