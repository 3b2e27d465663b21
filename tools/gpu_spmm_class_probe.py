"""Probe: one SpMM layer at C4 shape split by row class — the user rows (long: ~100 nonzeros) and the item rows (short: median ~3) —
each timed on its own for the library named by CR_LIB_PATH, to see whether a row-class-specific geometry would pay."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import coldrec_b200 as cr

dev = torch.device("cuda:0")
n_users, n_items, E, D = 1_000_000, 10_000_000, 100_000_000, 64
g = torch.Generator(device=dev).manual_seed(5)
wu = torch.exp(torch.randn(n_users, device=dev, generator=g))
wi = 1.0 / torch.arange(1, n_items + 1, device=dev, dtype=torch.float32) ** 0.8
wi = wi[torch.randperm(n_items, device=dev, generator=g)]
eu = torch.multinomial(wu, E, replacement=True, generator=g); ei = torch.multinomial(wi, E, replacement=True, generator=g)
G = cr.bipartite_norm_csr(eu, ei, n_users, n_items); del eu, ei, wu, wi
N = n_users + n_items
X = torch.randn(N, D, device=dev, generator=g) * 0.01
blocks = {"all": G, "users": G.row_block(0, n_users), "items": G.row_block(n_users, N)}


def ms_of(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


out = {"lib": os.path.basename(os.environ.get("CR_LIB_PATH", "default"))}
for name, B in blocks.items():
    Y = torch.empty(B.n_rows, D, device=dev)
    B.plan(D)
    t = ms_of(lambda: B.spmm(X, Y=Y))
    byts = B.nnz * (8 + 4 * D) + B.n_rows * 4 * D
    out[name] = {"ms": round(t, 3), "nnz": B.nnz, "rows": B.n_rows, "gbs": round(byts / t / 1e6, 1)}
print(json.dumps(out))
