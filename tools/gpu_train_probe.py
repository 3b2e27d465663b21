"""GPU probe: wall time of one BPR training step (forward + BPR + backward + Adam) and of the sampler, at C4 and C2 scale."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import coldrec_b200 as cr

dev = torch.device("cuda:0")


def timed(fn, iters, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def case(name, n_users, n_items, edges, bs, layers=3, d=64, iters=5):
    g = torch.Generator(device=dev).manual_seed(5)
    wu = torch.exp(torch.randn(n_users, device=dev, generator=g))
    wi = 1.0 / torch.arange(1, n_items + 1, device=dev, dtype=torch.float32) ** 0.8
    wi = wi[torch.randperm(n_items, device=dev, generator=g)]
    eu = torch.multinomial(wu, edges, replacement=True, generator=g)
    ei = torch.multinomial(wi, edges, replacement=True, generator=g)
    G = cr.bipartite_norm_csr(eu, ei, n_users, n_items)
    G.plan(d)
    N = n_users + n_items
    b = (6.0 / (N + 64)) ** 0.5
    U = (torch.rand(n_users, d, device=dev, generator=g) * 2 - 1) * b
    I = (torch.rand(n_items, d, device=dev, generator=g) * 2 - 1) * b
    smp = cr.PairwiseSampler(eu, ei, n_users, n_items, seed=1)
    step = cr.BprTrainStep(G, U, I, layers, 1e-3, 1e-4)
    out3 = torch.empty((3, bs), dtype=torch.int32, device=dev)
    t_sample = timed(lambda: smp.batch(0, 0, bs, out=out3), 20)
    u, i, j = smp.batch(0, 0, bs, out=out3)
    t_prop = timed(lambda: step._forward(), iters)
    t_grad = timed(lambda: step.gradients(u, i, j), iters)
    t_step = timed(lambda: step.step(u, i, j), iters)
    cap = step.capture(bs)
    def replay():
        smp.batch(0, 0, bs, out=cap.idx)
        cap()
    t_graph = timed(replay, iters * 4)
    nbytes_adam = N * d * 4 * 7
    print(json.dumps(dict(case=name, nnz=G.nnz, N=N, bs=bs, sample_ms=round(t_sample, 4), forward_ms=round(t_prop, 3), grad_ms=round(t_grad, 3),
                          step_ms=round(t_step, 3), graph_step_ms=round(t_graph, 3), adam_ms=round(t_step - t_grad, 3), adam_GBps=round(nbytes_adam / ((t_step - t_grad) * 1e-3) / 1e9, 1),
                          loss=step.loss.cpu().tolist(), exhausted=int(smp.n_exhausted.item()))), flush=True)
    # big-batch sampler throughput
    nb = min(smp.n_pairs, 1 << 24)
    o = torch.empty((3, nb), dtype=torch.int32, device=dev)
    t = timed(lambda: smp.batch(1, 0, nb, out=o), 5)
    print(json.dumps(dict(case=name, sampler_pairs=nb, ms=round(t, 3), Mpairs_per_s=round(nb / t / 1e3, 1))), flush=True)


case("C2 citeulike-shaped", 5551, 16980, 204986 * 8 // 10, 4096, iters=20)
case("C4 100M-edge", 1_000_000, 10_000_000, 100_000_000, 4096)
