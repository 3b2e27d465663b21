"""Small-shape drive of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool memcheck  python tools/gpu_sanitize.py [sweep|spmm|train|towers|all]
    compute-sanitizer --tool racecheck python tools/gpu_sanitize.py sweep

Shapes are tiny (the tools slow kernels down 10-100x) but reach every warp role of the tcgen05 sweep (TMA producer, MMA
issuer, bitmap producer, epilogue fast + slow path, buffer compaction), the seed phase on and off, the flag and compacted
(item_gids) mask paths, the grouped SpMM with long-row chunks riding in the same launch, and the BPR backward scatter.
Results are checked against torch on the device so a silent corruption would also show up here."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from coldrec_b200 import ops

dev = torch.device("cuda:0")
what = sys.argv[1] if len(sys.argv) > 1 else "all"
g = torch.Generator(device=dev).manual_seed(11)


def ref_topk(U, I, uids, rowptr, col, K, flags=None, excl=0):
    S = U[uids.long()] @ I.T
    rp = rowptr.tolist()
    for j in range(len(rp) - 1):
        S[j, col[rp[j]:rp[j + 1]].long()] = -1e9
    if flags is not None and excl:
        S[:, (flags & excl) != 0] = -1e9
    return torch.topk(S, K, dim=1)


def sweep_case(n_users, n_items, n_q, mask_per, seed_tiles, with_flags=False, compact=False, d=64):
    os.environ["CR_TC_SEED_TILES"] = str(seed_tiles)
    U = torch.randn(n_users, d, device=dev, generator=g) * 0.125
    I = torch.randn(n_items, d, device=dev, generator=g) * 0.125
    uids = torch.randperm(n_users, device=dev, generator=g)[:n_q].to(torch.int32)
    x = torch.sort(torch.randint(0, n_items - mask_per, (n_q, mask_per), device=dev, generator=g), dim=1).values
    col = (x + torch.arange(mask_per, device=dev)).to(torch.int32).flatten().contiguous()
    rowptr = torch.arange(0, (n_q + 1) * mask_per, mask_per, device=dev, dtype=torch.int64)
    flags = (torch.rand(n_items, device=dev, generator=g) < 0.2).to(torch.uint8) if (with_flags or compact) else None
    if compact:
        gids = torch.nonzero(flags == 0).flatten().to(torch.int32)
        s, i, nref = ops.score_topk(U, ops.gather_rows(I, gids), 20, user_ids=uids, item_gids=gids, mask_rowptr=rowptr, mask_col=col,
                                    precision=ops.SCORE_TF32_CHECKED)
    else:
        s, i, nref = ops.score_topk(U, I, 20, user_ids=uids, mask_rowptr=rowptr, mask_col=col, item_flags=flags,
                                    flag_exclude=1 if with_flags else 0, precision=ops.SCORE_TF32_CHECKED)
    rs, ri = ref_topk(U, I, uids, rowptr, col, 20, flags, 1 if (with_flags or compact) else 0)
    torch.cuda.synchronize()
    assert torch.allclose(s, rs, atol=1e-5), "scores differ"
    same = (i.long() == ri).float().mean().item()
    assert same > 0.999, f"ids differ ({same})"
    print(f"sweep d={d} n_q={n_q} n_items={n_items} seed={seed_tiles} flags={with_flags} compact={compact}: ok, refined {int(nref.item())}")


if what in ("sweep", "all"):
    sweep_case(600, 96 * 70, 300, 12, 0)                     # no seed phase, one partial unit, slow path flooded
    sweep_case(600, 96 * 1100 + 17, 520, 20, 128)            # seed phase on, three units, ragged last tile
    sweep_case(400, 96 * 80, 256, 8, 0, with_flags=True)     # in-kernel flag mask
    sweep_case(400, 96 * 90, 200, 8, 0, compact=True)        # compacted table: per-entry binary search in the bitmap producer
    sweep_case(500, 64 * 90 + 5, 300, 10, 0, d=128)          # d = 128 instantiation (64-item tiles, four TMA boxes per tile)
    sweep_case(500, 64 * 1100, 260, 10, 128, with_flags=True, d=128)
    sweep_case(300, 96 * 60 + 11, 200, 8, 0, d=48)            # narrower table: zero-padded copy in the workspace, d = 64 instantiation
    s, i, _ = ops.score_topk(torch.randn(64, 64, device=dev), torch.randn(5000, 64, device=dev), 20, precision=ops.SCORE_EXACT_F32)
    torch.cuda.synchronize()
    print("exact scorer: ok")

if what in ("spmm", "all"):
    from coldrec_b200 import CsrGraph
    n, d = 9000, 64
    deg = torch.randint(0, 12, (n,), device=dev, generator=g)
    deg[5], deg[77], deg[n - 1] = 3000, 700, 65                      # long rows: chunks ride in the first CTAs of the launch
    rowptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    rowptr[1:] = torch.cumsum(deg, 0)
    nnz = int(rowptr[-1])
    col = torch.randint(0, n, (nnz,), device=dev, generator=g).to(torch.int32)
    val = torch.randn(nnz, device=dev, generator=g)
    X = torch.randn(n, d, device=dev, generator=g)
    G = CsrGraph(rowptr, col, val, n)
    y = G.spmm(X, Y=torch.empty(n, d, device=dev))
    rows = torch.repeat_interleave(torch.arange(n, device=dev), deg)
    ref = torch.zeros(n, d, device=dev).index_add_(0, rows, X[col.long()] * val[:, None])
    torch.cuda.synchronize()
    assert (y - ref).abs().max().item() <= 1e-4 * ref.abs().max().item()
    for dd in (32, 128, 48):
        Xd = torch.randn(n, dd, device=dev, generator=g)
        G.spmm(Xd, Y=torch.empty(n, dd, device=dev))
    torch.cuda.synchronize()
    print("spmm (grouped rows + fused long-row chunks + reduce; d=64/32/128/48): ok")
    # the multi-GPU epilogue on one GPU: this GPU is its own (only) peer — shared-memory staging + TMA bulk stores, both geometries
    from coldrec_b200 import ops
    for force_big in (False, True):
        if force_big:
            os.environ["CR_SPMM_FORCE_BIG"] = "1"
        table = torch.full((n + 16, d), float("nan"), device=dev)
        ptrs = torch.tensor([table.data_ptr()], dtype=torch.int64, device=dev)
        need = (torch.rand(n, device=dev, generator=g) < 0.7).to(torch.uint8)
        acc = torch.zeros(n, d, device=dev)
        ops.spmm_bcast(rowptr, col, val, X, ptrs.data_ptr(), 1, 8, acc=acc, acc_beta=0.0, plan=G.plan(d), bcast_acc=True,
                       peer_row_split=5000, peer_row_offset_hi=16, peer_need=need)
        torch.cuda.synchronize()
        os.environ.pop("CR_SPMM_FORCE_BIG", None)
        got_lo, got_hi = table[8:8 + 5000], table[16 + 5000:16 + n]
        sent = need.bool()
        long_rows = deg > (512 if force_big else 64)
        assert (acc - ref).abs().max().item() <= 1e-4 * ref.abs().max().item()
        both = torch.cat([got_lo, got_hi])
        assert torch.equal(both[sent], acc[sent]), "staged rows differ from the local result"
        assert torch.isnan(both[~sent]).all(), "a row was delivered to a GPU that does not read it"
    print("spmm staged peer stores (TMA bulk copies, two destination ranges, need mask; small + big geometry): ok")

if what in ("train", "all"):
    import coldrec_b200 as cr
    nu, ni = 500, 3000
    eu = torch.randint(0, nu, (20000,), device=dev, generator=g)
    ei = torch.randint(0, ni, (20000,), device=dev, generator=g)
    G = cr.bipartite_norm_csr(eu, ei, nu, ni)
    E0u, E0i = torch.randn(nu, 64, device=dev, generator=g) * 0.1, torch.randn(ni, 64, device=dev, generator=g) * 0.1
    smp = cr.PairwiseSampler(eu.to(torch.int32), ei.to(torch.int32), nu, ni, seed=3)
    step = cr.BprTrainStep(G, E0u, E0i, 2, 1e-3, 1e-4)
    for k in range(2):
        u, i, j = smp.batch(0, k * 512, 512)
        loss = step.step(u, i, j)
    torch.cuda.synchronize()
    print("train step (sampler + propagate + bpr fwd/bwd + propagate + adam): ok", loss.cpu().tolist()[:2])

if what in ("towers", "all"):
    X = torch.randn(700, 300, device=dev, generator=g)
    X2 = torch.randn(700, 64, device=dev, generator=g)
    W = torch.randn(200, 364, device=dev, generator=g) * 0.05
    b = torch.randn(200, device=dev, generator=g)
    y = ops.linear_act(X2, W, b, X2=X, act="tanh")
    ref = torch.tanh(torch.cat([X2, X], 1) @ W.T + b)
    torch.cuda.synchronize()
    assert (y - ref).abs().max().item() < 1e-4
    y3, sp = ops.linear_act_tc(ops.split_tf32(X2), ops.split_tf32(W), b, X2=ops.split_tf32(X), act="tanh", want_split=True)
    torch.cuda.synchronize()
    assert (y3 - ref).abs().max().item() < 1e-5
    y4, _ = ops.linear_act_tc(X2, ops.split_tf32(W), b, X2=X, act="tanh")                 # raw rows: hi / lo split inside the kernel
    torch.cuda.synchronize()
    assert torch.equal(y4, y3)
    print("towers (SIMT + tcgen05 3xTF32): ok")
