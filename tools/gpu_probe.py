"""Scratch GPU probe (not part of the product): times the kernels at realistic sizes and prints JSON lines."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import coldrec_b200 as cr
from coldrec_b200 import ops

dev = torch.device("cuda:0")


def timed(fn, iters=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), ts


def synth_graph(n_users, n_items, n_edges, seed):
    g = torch.Generator(device=dev).manual_seed(seed)
    wu = torch.exp(torch.randn(n_users, device=dev, generator=g))
    wi = 1.0 / torch.arange(1, n_items + 1, device=dev, dtype=torch.float32) ** 0.8
    wi = wi[torch.randperm(n_items, device=dev, generator=g)]
    u = torch.multinomial(wu, n_edges, replacement=True, generator=g)
    i = torch.multinomial(wi, n_edges, replacement=True, generator=g)
    return u, i


def probe_spmm(n_users, n_items, n_edges, L=3):
    t0 = time.time()
    u, i = synth_graph(n_users, n_items, n_edges, 5)
    G = cr.bipartite_norm_csr(u, i, n_users, n_items)
    del u, i
    torch.cuda.synchronize()
    build_s = time.time() - t0
    N, d = n_users + n_items, 64
    b = (6.0 / (N + 64)) ** 0.5
    E0u = (torch.rand(n_users, d, device=dev) * 2 - 1) * b
    E0i = (torch.rand(n_items, d, device=dev) * 2 - 1) * b
    G.plan(d)
    deg = (G.rowptr[1:] - G.rowptr[:-1])
    ms, all_ms = timed(lambda: cr.propagate(G, E0u, E0i, L), iters=3)
    nnz = G.nnz
    bytes_layer = nnz * (8 + 4 * d) + 3 * N * 4 * d + 8 * (N + 1)
    X = torch.cat([E0u, E0i]).contiguous(); Y = torch.empty_like(X)
    ms1, _ = timed(lambda: G.spmm(X, Y=Y), iters=5)
    out = dict(kind="spmm", n_users=n_users, n_items=n_items, nnz=nnz, max_deg=int(deg.max()), long_rows=int((deg > 512).sum()),
               build_s=round(build_s, 2), propagate_ms=ms, all=all_ms, edges_per_s=nnz * L / (ms * 1e-3), gbps_alg=bytes_layer * L / (ms * 1e-3) / 1e9,
               single_spmm_ms=ms1, single_gbps=(nnz * (8 + 4 * d) + N * 4 * d + 8 * (N + 1)) / (ms1 * 1e-3) / 1e9)
    print(json.dumps(out), flush=True)
    return G


def probe_score(n_users, n_items, n_q, mask_per_user, precisions=(1, 0), K=20):
    g = torch.Generator(device=dev).manual_seed(6)
    U = torch.randn(n_users, 64, device=dev, generator=g) * 0.125
    I = torch.randn(n_items, 64, device=dev, generator=g) * 0.125
    uids = torch.randperm(n_users, device=dev, generator=g)[:n_q].to(torch.int32)
    rowptr = torch.arange(0, (n_q + 1) * mask_per_user, mask_per_user, device=dev, dtype=torch.int64)
    col = torch.sort(torch.randint(0, n_items, (n_q, mask_per_user), device=dev, generator=g), dim=1).values.to(torch.int32).flatten()
    res = {}
    for prec in precisions:
        need = ops._lib.load().cr_score_topk_workspace_bytes(n_q, n_items, 64, K, prec)
        ws = torch.empty(max(need, 1), dtype=torch.uint8, device=dev)
        fn = lambda: ops.score_topk(U, I, K, user_ids=uids, mask_rowptr=rowptr, mask_col=col, precision=prec, workspace=ws)
        ms, all_ms = timed(fn, iters=3)
        s, i, nref = fn()
        res[prec] = (s, i)
        flops = 2.0 * n_q * n_items * 64
        print(json.dumps(dict(kind="score", precision=prec, n_q=n_q, n_items=n_items, ms=ms, all=all_ms, users_per_s=n_q / (ms * 1e-3),
                              tflops=flops / (ms * 1e-3) / 1e12, n_refined=int(nref.item()), ws_mb=need / 2**20)), flush=True)
    if len(res) == 2:
        (s1, i1), (s0, i0) = res[1], res[0]
        print(json.dumps(dict(kind="score_agree", ids_equal=float((i1 == i0).float().mean()), max_score_diff=float((s1 - s0).abs().max()))), flush=True)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    print(json.dumps(dict(gpu=torch.cuda.get_device_name(0))), flush=True)
    if which in ("all", "score"):
        probe_score(100000, 200000, 4096, 100)
        probe_score(200000, 2000000, 16384, 100)
        probe_score(1000000, 10000000, 65536, 100, precisions=(1,))
    if which in ("all", "spmm"):
        probe_spmm(200000, 2000000, 20000000)
        probe_spmm(1000000, 10000000, 100000000)
