"""GPU probe: one LightGCN training step at CiteULike scale (launch list under ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import coldrec_b200 as cr
dev = torch.device("cuda:0")
n_users, n_items, edges, bs, d = 5551, 16980, 163988, 4096, 64
g = torch.Generator(device=dev).manual_seed(5)
wu = torch.exp(torch.randn(n_users, device=dev, generator=g))
wi = 1.0 / torch.arange(1, n_items + 1, device=dev, dtype=torch.float32) ** 0.8
eu = torch.multinomial(wu, edges, replacement=True, generator=g)
ei = torch.multinomial(wi, edges, replacement=True, generator=g)
G = cr.bipartite_norm_csr(eu, ei, n_users, n_items); G.plan(d)
U = torch.randn(n_users, d, device=dev) * 0.1; I = torch.randn(n_items, d, device=dev) * 0.1
smp = cr.PairwiseSampler(eu, ei, n_users, n_items, seed=1)
step = cr.BprTrainStep(G, U, I, 3, 1e-3, 1e-4)
for k in range(3):
    u, i, j = smp.batch(0, k * bs, bs)
    step.step(u, i, j)
torch.cuda.synchronize()
print("max row nnz", int((G.rowptr[1:] - G.rowptr[:-1]).max()), "nnz", G.nnz)
