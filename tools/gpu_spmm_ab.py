"""A/B probe for K3 at C4 size: one SpMM (Y only) and the 3-layer propagation, for the library named by CR_LIB_PATH."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import coldrec_b200 as cr

dev = torch.device("cuda:0")
n_users, n_items, n_edges, d = 1_000_000, 10_000_000, int(os.environ.get("AB_EDGES", 100_000_000)), 64
g = torch.Generator(device=dev).manual_seed(5)
wu = torch.exp(torch.randn(n_users, device=dev, generator=g))
wi = 1.0 / torch.arange(1, n_items + 1, device=dev, dtype=torch.float32) ** 0.8
wi = wi[torch.randperm(n_items, device=dev, generator=g)]
eu = torch.multinomial(wu, n_edges, replacement=True, generator=g)
ei = torch.multinomial(wi, n_edges, replacement=True, generator=g)
G = cr.bipartite_norm_csr(eu, ei, n_users, n_items)
del eu, ei
N = n_users + n_items
b = (6.0 / (N + 64)) ** 0.5
E0u = (torch.rand(n_users, d, device=dev, generator=g) * 2 - 1) * b
E0i = (torch.rand(n_items, d, device=dev, generator=g) * 2 - 1) * b
X = torch.cat([E0u, E0i]); Y = torch.empty_like(X)
G.plan(d)


def timed(fn, iters=6, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b_.record(); torch.cuda.synchronize()
    return a.elapsed_time(b_) / iters


ms1 = timed(lambda: G.spmm(X, Y=Y))
ms3 = timed(lambda: cr.propagate(G, E0u, E0i, 3))
chk = float(Y.double().abs().sum())
bytes_layer = G.nnz * (8 + 4 * d) + 3 * N * 4 * d + 8 * (N + 1)
print(json.dumps(dict(lib=os.path.basename(os.environ.get("CR_LIB_PATH", "default")), spmm_ms=round(ms1, 3), propagate3_ms=round(ms3, 3),
                      gbps_alg=round(bytes_layer * 3 / ms3 / 1e6, 1), checksum=chk)), flush=True)
