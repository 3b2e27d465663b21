"""The reference's own GPU path on this box — the "existing Blackwell library path" of SURVEY §8(d): the torch calls of
model/BaseRecommender.py:170-182 + model/MF.py:62 (cuBLAS SGEMM with TF32 off, index_put mask loop, ATen topk) and of
model/LightGCN.py:86-96 (torch.sparse.mm on the coalesced int64 COO tensor, stack + mean), timed with CUDA events on the
bench's C5 / C4 inputs.  Not part of the product; prints one JSON line per workload."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

dev = torch.device("cuda:0")
D, K = 64, 20


def timed(fn, iters, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def scoring(n_users=1_000_000, n_items=10_000_000, batch=512, mask_per=100):
    g = torch.Generator(device=dev).manual_seed(1)
    U = torch.randn(n_users, D, device=dev, generator=g) * 0.125
    I = torch.randn(n_items, D, device=dev, generator=g) * 0.125
    users = torch.arange(batch, device=dev)
    x = torch.sort(torch.randint(0, n_items - mask_per, (batch, mask_per), device=dev, generator=g), dim=1).values
    rated = [(x[j] + torch.arange(mask_per, device=dev)) for j in range(batch)]       # per-user LongTensors, as _get_eval_cache builds

    def reference_batch():
        cand = torch.matmul(U[users], I.transpose(0, 1))            # MF.py:62
        for j in range(batch):                                      # BaseRecommender.py:175-177
            cand[j, rated[j]] = -10e8
        return torch.topk(cand, K, dim=1, largest=True, sorted=True)   # :182

    rows = torch.arange(batch, device=dev).repeat_interleave(mask_per)
    cols = torch.cat(rated)

    def vectorised_mask_batch():                                    # same library kernels, the Python loop replaced by one index_put
        cand = torch.matmul(U[users], I.transpose(0, 1))
        cand[rows, cols] = -10e8
        return torch.topk(cand, K, dim=1, largest=True, sorted=True)

    ms_ref = timed(reference_batch, 3)
    ms_vec = timed(vectorised_mask_batch, 3)
    ms_mm = timed(lambda: torch.matmul(U[users], I.transpose(0, 1)), 3)
    print(json.dumps(dict(workload=f"C5 scoring, torch GPU library path, batch {batch} users x {n_items} items (a 4096-user batch would be 164 GB)",
                          ms_per_batch=round(ms_ref, 2), users_per_s=round(batch / ms_ref * 1e3, 1),
                          vectorised_mask_users_per_s=round(batch / ms_vec * 1e3, 1), matmul_only_ms=round(ms_mm, 2),
                          matmul_tflops=round(2.0 * batch * n_items * D / ms_mm / 1e9, 1), allow_tf32=torch.backends.cuda.matmul.allow_tf32)), flush=True)


def propagation(n_users=1_000_000, n_items=10_000_000, n_edges=100_000_000, L=3):
    import coldrec_b200 as cr
    g = torch.Generator(device=dev).manual_seed(5)
    wu = torch.exp(torch.randn(n_users, device=dev, generator=g))
    wi = 1.0 / torch.arange(1, n_items + 1, device=dev, dtype=torch.float32) ** 0.8
    wi = wi[torch.randperm(n_items, device=dev, generator=g)]
    eu = torch.multinomial(wu, n_edges, replacement=True, generator=g)
    ei = torch.multinomial(wi, n_edges, replacement=True, generator=g)
    G = cr.bipartite_norm_csr(eu, ei, n_users, n_items)          # the builder only; the timed path below is torch's
    del eu, ei, wu, wi
    N = n_users + n_items
    rows = torch.repeat_interleave(torch.arange(N, device=dev), G.rowptr[1:] - G.rowptr[:-1])
    A = torch.sparse_coo_tensor(torch.stack([rows, G.col.long()]), G.val, (N, N)).coalesce()      # databuilder.py:959-962
    nnz = G.nnz
    del rows, G
    torch.cuda.empty_cache()
    b = (6.0 / (N + 64)) ** 0.5
    E0 = (torch.rand(N, D, device=dev, generator=g) * 2 - 1) * b

    def forward():                                                  # LightGCN.py:86-96
        ego, layers = E0, [E0]
        for _ in range(L):
            ego = torch.sparse.mm(A, ego)
            layers.append(ego)
        return torch.mean(torch.stack(layers, dim=1), dim=1)

    ms = timed(forward, 3)
    print(json.dumps(dict(workload=f"C4 LightGCN {L}-layer propagation, torch.sparse.mm (COO int64) + stack/mean on the GPU, nnz={nnz}",
                          ms_per_step=round(ms, 2), edges_per_s=round(nnz * L / ms * 1e3, 1))), flush=True)


if __name__ == "__main__":
    for fn in (scoring, propagation):
        try:
            fn()
        except Exception as ex:          # an OOM here must not take the box down with it
            print(json.dumps(dict(workload=fn.__name__, error=f"{type(ex).__name__}: {str(ex)[:200]}")), flush=True)
        torch.cuda.empty_cache()
