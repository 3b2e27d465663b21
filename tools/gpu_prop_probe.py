"""Probe (torchrun, N ranks): the C4 row-partitioned propagation, graph built once, every exchange variant timed in turn with
per-layer CUDA events (kernel time and barrier wait per layer, min / max over ranks) and checked against rank-local
single-GPU propagation.   python -m torch.distributed.run --nproc-per-node N tools/gpu_prop_probe.py [steps]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import coldrec_b200 as cr
from coldrec_b200.dist import RowPartitionedGraph

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"])); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
n_users, n_items, E, D, L = 1_000_000, 10_000_000, 100_000_000, 64, 3
g = torch.Generator(device=dev).manual_seed(5)
wu = torch.exp(torch.randn(n_users, device=dev, generator=g))
wi = 1.0 / torch.arange(1, n_items + 1, device=dev, dtype=torch.float32) ** 0.8
wi = wi[torch.randperm(n_items, device=dev, generator=g)]
eu = torch.multinomial(wu, E, replacement=True, generator=g); ei = torch.multinomial(wi, E, replacement=True, generator=g)
G = cr.bipartite_norm_csr(eu, ei, n_users, n_items); del eu, ei, wu, wi
nnz_t = torch.tensor([G.nnz], dtype=torch.int64, device=dev); dist.broadcast(nnz_t, 0); n0 = int(nnz_t.item())
rp = G.rowptr if rank == 0 else torch.empty(n_users + n_items + 1, dtype=torch.int64, device=dev)
col = G.col if rank == 0 else torch.empty(n0, dtype=torch.int32, device=dev)
val = G.val if rank == 0 else torch.empty(n0, dtype=torch.float32, device=dev)
for t_ in (rp, col, val): dist.broadcast(t_, 0)
G = cr.CsrGraph(rp, col, val, n_users + n_items)
N = n_users + n_items; b = (6.0 / (N + 64)) ** 0.5
torch.manual_seed(5)
E0 = (torch.rand(N, D, device=dev, generator=g) * 2 - 1) * b
dist.broadcast(E0, 0)
t0 = time.time()
PG = RowPartitionedGraph(G.rowptr.cpu().numpy(), G.col.cpu().numpy(), G.val.cpu().numpy(), dev, segments=(n_users, n_items))
t_part = time.time() - t0
PG.local.plan(D); PG.enable_p2p(D)
ru, ri = cr.propagate(G, E0[:n_users].contiguous(), E0[n_users:].contiguous(), L)
ref = torch.cat([ru, ri]); del ru, ri
scale = ref.abs().max().item()
# ---- what the fabric itself does with this exchange: NCCL all-gather, copy-engine pushes, the kernel's push path with no SpMM work
def timed(fn, iters=8):
    for _ in range(2): fn()
    dist.barrier(device_ids=[dev.index]); torch.cuda.synchronize()
    a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b_.record(); torch.cuda.synchronize()
    t_ = torch.tensor([a.elapsed_time(b_) / iters], dtype=torch.float64, device=dev)
    dist.all_reduce(t_, op=dist.ReduceOp.MAX)
    return round(float(t_.item()), 3)
from coldrec_b200 import ops
nl, rp_ = PG.n_local, PG.rows_pad
y = torch.randn(rp_, D, device=dev)
ag_out = torch.empty(world * rp_, D, device=dev)
fabric = {"bytes_per_gpu_egress": nl * D * 4 * (world - 1)}
fabric["nccl_all_gather_ms"] = timed(lambda: dist.all_gather_into_tensor(ag_out, y))
try:
    h = PG._xhdl[2]
    peers = [h.get_buffer(p, (world * rp_, D), torch.float32) for p in range(world)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(world)]
    def ce_push():
        cur = torch.cuda.current_stream()
        for k in range(1, world):
            p = (rank + k) % world
            streams[p].wait_stream(cur)
            with torch.cuda.stream(streams[p]):
                peers[p][rank * rp_:rank * rp_ + nl].copy_(y[:nl], non_blocking=True)
        for k in range(1, world):
            cur.wait_stream(streams[(rank + k) % world])
        h.barrier()
    fabric["copy_engine_push_ms"] = timed(ce_push)
except Exception as ex:
    fabric["copy_engine_push_ms"] = f"{type(ex).__name__}: {str(ex)[:100]}"
zrp = torch.zeros(nl + 1, dtype=torch.int64, device=dev); zc = torch.zeros(1, dtype=torch.int32, device=dev)[:0]
acc0 = torch.zeros(rp_, D, device=dev)
def kernel_push():
    ops.spmm_bcast(zrp, zc, None, E0, PG._xhdl[2].buffer_ptrs_dev, world, rank * rp_, acc=acc0[:nl], acc_beta=1.0, acc_div=1.0,
                   plan=None, bcast_acc=True)
    PG._xhdl[2].barrier()
for nm, env in (("kernel_push_tma_ms", {}), ("kernel_push_st_ms", {"CR_SPMM_PEER_ST": "1"})):
    os.environ.pop("CR_SPMM_PEER_ST", None); os.environ.update(env)
    try:
        fabric[nm] = timed(kernel_push)
    except Exception as ex:
        fabric[nm] = f"{type(ex).__name__}: {str(ex)[:100]}"
os.environ.pop("CR_SPMM_PEER_ST", None)
if rank == 0:
    print(json.dumps({"fabric": fabric, "world": world, "what": "dense exchange of one layer (every local row to every peer), no SpMM work"}), flush=True)

variants = [("tma full", {}, {}), ("tma full, fixed 128 rows per warp", {"CR_SPMM_FIXED_ROWS": "1"}, {}), ("st full", {"CR_SPMM_PEER_ST": "1"}, {}),
            ("tma users", {}, {"replicate_result": (0,)}), ("tma dense", {}, {"sparse": False})]
for name, env, kw in variants:
    os.environ.pop("CR_SPMM_PEER_ST", None); os.environ.pop("CR_SPMM_FIXED_ROWS", None); os.environ.update(env)
    run = lambda ev=None: PG.propagate_p2p(E0, L, copy=False, layer_events=ev, **kw)
    for _ in range(3): res = run()
    if "replicate_result" in kw:
        (ub, ue), (ib, ie) = PG.parts[rank]
        err = max((res[:n_users] - ref[:n_users]).abs().max().item(), (res[ib:ie] - ref[ib:ie]).abs().max().item()) / scale
    else:
        err = (res - ref).abs().max().item() / scale
    dist.barrier(device_ids=[dev.index]); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): run()
    e1.record(); dist.barrier(device_ids=[dev.index]); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    evs = []
    for _ in range(4): run(evs)
    torch.cuda.synchronize()
    k = np.array([[e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])] for e in evs]).reshape(4, L, 2).mean(0)   # [layer, (kernel, barrier)]
    t_ = torch.tensor([ms, err] + k.flatten().tolist(), dtype=torch.float64, device=dev)
    mx, mn = t_.clone(), t_.clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX); dist.all_reduce(mn, op=dist.ReduceOp.MIN)
    if rank == 0:
        mx, mn = mx.tolist(), mn.tolist()
        print(json.dumps({"variant": name, "world": world, "ms_per_step": round(mx[0], 3), "rel_err_max": mx[1],
                          "layer_kernel_ms_max": [round(mx[2 + 2 * l], 3) for l in range(L)], "layer_kernel_ms_min": [round(mn[2 + 2 * l], 3) for l in range(L)],
                          "layer_barrier_ms_max": [round(mx[3 + 2 * l], 3) for l in range(L)], "layer_barrier_ms_min": [round(mn[3 + 2 * l], 3) for l in range(L)],
                          "partition_s": round(t_part, 1), "need_copies": round(PG.need_copies, 2), "n_local_rows": PG.n_local, "nnz_local": PG.local.nnz}), flush=True)
dist.destroy_process_group()
