#!/usr/bin/env python
"""bench.py — full-rank users/sec (top-20 over the catalogue) and LightGCN edges/sec on B200.

    python bench.py --gpus 1 --steps 8 --warmup 3                       # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...   # N GPUs, one rank each
    python bench.py --impl reference ...                                # the reference's CPU path (oracle port), rank 0

Workload (BASELINE.json configs[4] and configs[3], synthetic, seeded):
  scoring   1M users x 10M items, d=64, K=20, ~100 train-masked items and 10 ground-truth items per user.
            A step = one eval batch of 75,776 x N users against the whole catalogue: fused score + train mask +
            top-20 + Hit/Precision/Recall/NDCG@{10,20} reduction.  N GPUs shard the catalogue (10M/N items each),
            exchange (score,id) candidates over NCCL and merge; per-GPU work is constant -> weak scaling.
  lightgcn  1M users + 10M items, 100M interactions (nnz(A) ~ 2e8), 3-layer propagation with fused layer mean.
Inputs are larger than L2 (item shard >= 320 MB, embedding table 2.8 GB), so no explicit L2 flush is needed.
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_USERS, N_ITEMS, D, K, TOPN = 1_000_000, 10_000_000, 64, 20, [10, 20]
USERS_PER_STEP = 75_776          # 296 query tiles of 256 = two full waves of 148 SMs
MASK_PER_USER, GT_PER_USER = 100, 10
MAX_DISTINCT_PLANS = 12           # distinct synthetic eval batches kept resident (device + pinned host); longer runs cycle through them
GRAPH_EDGES, LAYERS = 100_000_000, 3


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="both", choices=["both", "score", "lightgcn"])
    ap.add_argument("--users-per-step", type=int, default=USERS_PER_STEP)
    ap.add_argument("--n-items", type=int, default=N_ITEMS)
    ap.add_argument("--n-users", type=int, default=N_USERS)
    ap.add_argument("--graph-edges", type=int, default=GRAPH_EDGES)
    ap.add_argument("--cpu-sample-users", type=int, default=128)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-d128", action="store_true", help="skip the d = 128 (VBPR / AMR two-product) scoring line (N=1; under \"d128\")")
    ap.add_argument("--no-robustness", action="store_true", help="skip the robustness lines (N=1: warm/cold settings, duplicate rows, "
                                                                 "heavy-tailed / norm-sorted item tables; under \"robustness\")")
    ap.add_argument("--configs", default="C1,C2,C3", help="dataset-shaped side lines under \"extra\" (N=1 only): any of C1,C2,C3")
    ap.add_argument("--no-configs", action="store_true", help="skip the C1 / C2 / C3 dataset-shaped lines")
    ap.add_argument("--config-steps", type=int, default=5)
    ap.add_argument("--shard", default="auto", choices=["auto", "items", "users"],
                    help="N>1 scoring layout.  items: the catalogue split N ways + NCCL candidate all-gather (north star).  auto "
                         "(default): the same, but item shards are kept at >= --min-shard-items items; beyond that the ranks form "
                         "user groups (N=8: two 5M-item shards x four user groups).  users: item table replicated, no exchange")
    ap.add_argument("--min-shard-items", type=int, default=5_000_000,
                    help="--shard auto keeps item shards at least this long (the sweep runs at 0.94 of the tensor peak on 5M-item "
                         "shards, 0.90 on 2.5M, 0.81 on 1.25M): N = 2, 4, 8 -> 2 item shards x N/2 user groups")
    ap.add_argument("--prop-result", default="full", choices=["full", "users"],
                    help="multi-GPU propagation result: the whole table on every GPU (as at N=1), or user rows replicated + item rows "
                         "with their owner (what item-sharded scoring consumes)")
    ap.add_argument("--multicast", action="store_true",
                    help="multi-GPU propagation: rows wanted by every GPU leave as one NVLS multimem.st (measured slower; default unicast)")
    ap.add_argument("--peer-store", default="tma", choices=["tma", "st"],
                    help="multi-GPU propagation: how finished rows reach the peers — staged in shared memory and pushed by TMA bulk "
                         "copies (default), or one 16-byte st.global per lane and destination (round-1 path, kept for A/B)")
    ap.add_argument("--dense-exchange", action="store_true", help="multi-GPU propagation: send every row to every GPU (no need masks)")
    ap.add_argument("--no-train", action="store_true", help="skip the LightGCN training-step line of the lightgcn workload")
    ap.add_argument("--train-batch", type=int, default=4096)
    return ap.parse_args()


# ------------------------------------------------------------------------------------------- helpers
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING a timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.index, self.lines, self.proc, self.tail = index, [], None, 0

    def top_up(self, run, min_samples=3, max_s=3.0):
        """A timed region shorter than nvidia-smi's start-up + period (K steps of ~10 ms) ends before the first sample:
        keep the SAME step running, untimed, until a few samples under that load exist; their number is reported."""
        if not self.proc:
            return
        n0, t0 = len(self.lines), time.time()
        while len(self.lines) < max(n0, 0) + (min_samples if n0 < min_samples else 0) and time.time() - t0 < max_s:
            run()
            torch.cuda.synchronize()
        self.tail = len(self.lines) - n0

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True)
            self.th.start()
        except OSError:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.th.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = max(mx, float(f[1]))
            except ValueError:
                continue
            for name, v in zip(self.NAMES, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [x for x in sm if x > 0.5 * max(sm)] if sm else []
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm), "samples_after_timed_region_same_load": self.tail}


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except OSError:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def ncu_traffic(key, src):
    """DRAM bytes per launch of the dominant kernel (dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full`
    capture, profiles/r02_traffic.json, written by tools/ncu_traffic.py).  DRAM counters cannot be read without a profiler,
    so the figure is a committed measurement — but one pinned to the kernel source it was taken on: the file records the
    sha256 of coldrec_b200/csrc/<src>, and when that no longer matches the tree the figure is withheld (null) and the
    reason reported, instead of going stale silently.  Returns (bytes | None, note | None)."""
    import hashlib
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json"))).get(key)
    except OSError:
        return None, "no committed ncu capture (profiles/r02_traffic.json)"
    if not rec:
        return None, f"no committed ncu capture of {key}"
    sha = hashlib.sha256(open(os.path.join(ROOT, "coldrec_b200", "csrc", src), "rb").read()).hexdigest()
    if rec.get("source_sha256") != sha:
        return None, f"stale: the committed capture was taken on another revision of csrc/{src} — re-capture (tools/gpu_round_check.sh + tools/ncu_traffic.py)"
    return rec["bytes_per_launch"], None


def make_step_plans(n_steps, n_q, n_users, n_items, seed, device):
    """Per step: eval user ids, train-mask CSR (MASK_PER_USER sorted item ids per user), ground-truth CSR."""
    g = torch.Generator(device=device).manual_seed(seed)
    plans = []
    mrp = torch.arange(0, (n_q + 1) * MASK_PER_USER, MASK_PER_USER, device=device, dtype=torch.int64)
    grp = torch.arange(0, (n_q + 1) * GT_PER_USER, GT_PER_USER, device=device, dtype=torch.int64)
    for s in range(n_steps):
        uids = ((torch.arange(n_q, device=device, dtype=torch.int64) + s * n_q) % n_users).to(torch.int32)
        # distinct sorted ids per row: sort random draws and nudge duplicates apart
        def rows(per):
            x = torch.sort(torch.randint(0, n_items - per, (n_q, per), device=device, generator=g), dim=1).values
            return (x + torch.arange(per, device=device)).to(torch.int32).flatten().contiguous()
        plans.append(dict(user_ids=uids, mask_rowptr=mrp, mask_col=rows(MASK_PER_USER), gt_rowptr=grp, gt_col=rows(GT_PER_USER)))
    return plans


def host_threads():
    """Host threads this process may use (torchrun exports OMP_NUM_THREADS=1: the CPU arms override it explicitly)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def score_workload(args, n_q):
    return (f"C5 full-catalog top-{K} scoring: {args.n_users} users x {args.n_items} items, d={D}, {n_q} users/step, "
            f"~{MASK_PER_USER} train-masked + {GT_PER_USER} gt items/user, Recall/NDCG@{TOPN} on device")


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def max_over_ranks(x, device, world):
    if world == 1:
        return x
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier(world, device):
    if world > 1:
        import torch.distributed as dist
        dist.barrier(device_ids=[device.index])
    torch.cuda.synchronize(device)


# ------------------------------------------------------------------------------------------- B200 arm
def run_b200(args):
    import coldrec_b200 as cr
    from coldrec_b200 import _lib, ops
    from coldrec_b200.dist import GridShardedFullRankScorer, grid_item_shards, shard_range
    from coldrec_b200.scoring import EvalPlan, HostBatchEvaluator

    rank, local_rank, world = dist_env()
    if args.gpus != world:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch N>1 with torch.distributed.run (one rank per GPU)")
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)
    lib = _lib.load()
    pk, pk_src = peaks()
    W, Ksteps = args.warmup, args.steps
    out = {}

    if args.workload in ("both", "score"):
        n_q = args.users_per_step * world
        g = torch.Generator(device=device).manual_seed(1)
        user_tab = torch.randn(args.n_users, D, device=device, generator=g) * 0.125
        S = {"items": world, "users": 1, "auto": grid_item_shards(world, args.n_items, args.min_shard_items)}[args.shard]
        scorer = GridShardedFullRankScorer(K, S, ops.SCORE_TF32_CHECKED)
        ib, ie = scorer.item_range(args.n_items)
        gi = torch.Generator(device=device).manual_seed(1000 + scorer.ishard)    # replicas of a shard hold the same rows
        item_shard = torch.randn(ie - ib, D, device=device, generator=gi) * 0.125
        # every step ranks a different batch of users; beyond MAX_DISTINCT_PLANS the batches repeat (a long --steps run would
        # otherwise pin W+K x 35 MB x N of host memory per rank) — each step still copies and sweeps its whole plan
        n_distinct = min(W + Ksteps, MAX_DISTINCT_PLANS)
        distinct = make_step_plans(n_distinct, n_q, args.n_users, args.n_items, 6, device)
        plans_d = [distinct[k % n_distinct] for k in range(W + Ksteps)]
        distinct_plans = [EvalPlan.from_arrays(**p) for p in distinct]
        plans = [distinct_plans[k % n_distinct] for k in range(W + Ksteps)]

        def step(plan):
            s, i = scorer.topk(user_tab, item_shard, ib, plan)
            return s, i, scorer.metrics(i, plan, TOPN, rounded=False)

        for p in plans[:W]:
            step(p)
        barrier(world, device)
        launches0 = lib.cr_launch_count()
        lib.cr_profile_enable(1)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(device.index) as clk:
            ev0.record()
            for p in plans[W:]:
                _, _, perf = step(p)
            ev1.record()
            barrier(world, device)
        ms = max_over_ranks(ev0.elapsed_time(ev1), device, world)
        launches = lib.cr_launch_count() - launches0
        import ctypes
        tot, cnt = ctypes.c_double(), ctypes.c_int()
        lib.cr_profile_read(0, ctypes.byref(tot), ctypes.byref(cnt))
        lib.cr_profile_enable(0)
        value = n_q * Ksteps / (ms * 1e-3)
        sweep_ms = tot.value / max(cnt.value, 1)
        n_q_rank = (lambda r: r[1] - r[0])(scorer.group_slice(n_q))      # users this rank sweeps against its item shard
        flops = 2.0 * n_q_rank * (ie - ib) * D                  # algorithmic FLOPs of one sweep launch (this rank's shard)
        tf32_peak = pk["bf16_tflops_sustained"] / 2.0             # TF32 dense = half the bf16 rate; kernel runs inside a long step
        achieved = flops / (sweep_ms * 1e-3) / 1e12 if sweep_ms > 0 else 0.0
        roofline = {"bound": "tensor", "kernel": "score_sweep_tc_kernel", "achieved": round(achieved, 1), "peak": round(tf32_peak, 1),
                    "unit": "TFLOP/s", "frac": round(achieved / tf32_peak, 4),
                    "traffic": None,
                    "peak_source": f"{pk_src} bf16_tflops_sustained/2 (TF32 dense is half the bf16 rate)",
                    "launch_ms": round(sweep_ms, 3), "launches": cnt.value, "flop_per_launch": flops,
                    "share_of_step": round(sweep_ms * cnt.value / ms, 4)}
        if world == 1 and n_q == USERS_PER_STEP and args.n_items == N_ITEMS:
            roofline["traffic"], note = ncu_traffic("score_sweep_tc_kernel", "score_tc.cu")
            if note:
                roofline["traffic_note"] = note

        # end to end through the host-buffer API: H2D of the step's plan from pinned memory, D2H of top-K + metric sums
        hb = HostBatchEvaluator(scorer, TOPN, n_q, n_q * MASK_PER_USER, n_q * GT_PER_USER, device)
        pinned = [hb.pin({k: v.cpu() for k, v in p.items()}) for p in distinct]     # this rank's user group of every host plan
        host_plans = [pinned[k % n_distinct] for k in range(W, W + Ksteps)]
        hb.run(user_tab, item_shard, ib, host_plans[0])
        barrier(world, device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k, hp in enumerate(host_plans):      # every step copies its own plan; step k+1's copy overlaps step k's sweep
            res = hb.run(user_tab, item_shard, ib, hp, host_plans[k + 1] if k + 1 < len(host_plans) else None)
        e1.record()
        barrier(world, device)
        e2e_ms = max_over_ranks(e0.elapsed_time(e1), device, world)
        out.update(metric="full-rank users/sec (top-20 over catalog)", value=round(value, 1), unit="users/s", n_gpus=world,
                   steps=Ksteps, warmup=W, ms_per_step=round(ms / Ksteps, 3), higher_is_better=True, scaling="weak",
                   vs_baseline=None, dtype="tf32 select + f32 rescore", data="synthetic",
                   config={"workload": score_workload(args, n_q),
                           "parallelism": ("single GPU" if world == 1 else
                                           f"user-sharded x{world}, item table replicated, no candidate exchange" if S == 1 else
                                           f"item-sharded x{S} ({(ie - ib)} items per shard)"
                                           + (f" x {world // S} user groups" if S < world else "")
                                           + " + NCCL candidate all-gather" + (" inside each group" if S < world else "")),
                           "l2": "inputs larger than L2 (item shard %.0f MB); no flush" % ((ie - ib) * D * 4 / 2**20),
                           "users_per_step": n_q, "n_items": args.n_items, "K": K},
                   e2e={"value": round(n_q * Ksteps / (e2e_ms * 1e-3), 1), "unit": "users/s", "h2d_bytes_per_step": hb.h2d_bytes,
                        "d2h_bytes_per_step": hb.d2h_bytes, "ms_per_step": round(e2e_ms / Ksteps, 3)},
                   gpu_launches=int(launches), roofline=roofline, clocks=clk.summary(),
                   check={"ndcg@20": perf[1][3], "recall@20": perf[1][2],
                          "n_refined_last_step": int(scorer.last_n_refined.item()) if scorer.last_n_refined is not None else None})
        if world > 1:
            out["check"].update(check_sharded_ids(args, scorer, user_tab, item_shard, ib, plans[-1], S, device, world))
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_score_baseline(user_tab, item_shard, plans_d[W], args.cpu_sample_users, args)
            out["gpu_library_baseline"] = library_score_baseline(user_tab, item_shard, plans_d[W])
        if world == 1 and not args.no_robustness:
            out["robustness"] = run_robustness(args, lib, user_tab, item_shard, distinct_plans, device, pk)
        if world == 1 and not args.no_d128:
            del item_shard
            item_shard = None
            torch.cuda.empty_cache()
            out["d128"] = run_d128(args, lib, distinct_plans, device, pk)
        del item_shard, plans, plans_d, host_plans, hb, distinct, distinct_plans, pinned
        torch.cuda.empty_cache()

    if args.workload in ("both", "lightgcn"):
        lg = run_lightgcn(args, device, rank, world, pk, pk_src, lib)
        if args.workload == "lightgcn":
            out = lg
        else:
            out["lightgcn"] = lg
    if world == 1 and not args.no_configs:
        import bench_configs
        torch.cuda.empty_cache()
        out["extra"] = {}
        for key in [k for k in args.configs.split(",") if k in bench_configs.CONFIGS]:
            try:
                out["extra"][key] = bench_configs.run_config(key, device, lib, steps=args.config_steps, warmup=2,
                                                             cpu=not args.no_cpu_baseline, host_threads=host_threads())
            except Exception as ex:      # a dataset-shaped side line must never cost the headline line
                out["extra"][key] = {"error": f"{type(ex).__name__}: {str(ex)[:200]}"}
            torch.cuda.empty_cache()
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def run_lightgcn(args, device, rank, world, pk, pk_src, lib):
    import ctypes
    import coldrec_b200 as cr
    W, Ksteps = args.warmup, args.steps
    n_users, n_items = args.n_users, args.n_items
    g = torch.Generator(device=device).manual_seed(5)
    wu = torch.exp(torch.randn(n_users, device=device, generator=g))                       # lognormal user activity
    wi = 1.0 / torch.arange(1, n_items + 1, device=device, dtype=torch.float32) ** 0.8     # Zipf(0.8) item popularity
    wi = wi[torch.randperm(n_items, device=device, generator=g)]
    eu = torch.multinomial(wu, args.graph_edges, replacement=True, generator=g)
    ei = torch.multinomial(wi, args.graph_edges, replacement=True, generator=g)
    G = cr.bipartite_norm_csr(eu, ei, n_users, n_items)
    del eu, ei, wu, wi
    if world > 1:      # multinomial sampling is not bit-reproducible across devices: every rank takes rank 0's graph
        import torch.distributed as dist
        nnz_t = torch.tensor([G.nnz], dtype=torch.int64, device=device)
        dist.broadcast(nnz_t, 0)
        n0 = int(nnz_t.item())
        rp = G.rowptr if rank == 0 else torch.empty(n_users + n_items + 1, dtype=torch.int64, device=device)
        col = G.col if rank == 0 else torch.empty(n0, dtype=torch.int32, device=device)
        val = G.val if rank == 0 else torch.empty(n0, dtype=torch.float32, device=device)
        for t_ in (rp, col, val):
            dist.broadcast(t_, 0)
        G = cr.CsrGraph(rp, col, val, n_users + n_items)
        torch.manual_seed(5)
    N = n_users + n_items
    b = (6.0 / (N + 64)) ** 0.5
    E0u = (torch.rand(n_users, D, device=device, generator=g) * 2 - 1) * b
    E0i = (torch.rand(n_items, D, device=device, generator=g) * 2 - 1) * b
    exchange = None
    if world == 1:
        G.plan(D)
        bufs = cr.PropagationBuffers(n_users + n_items, D, device, with_ego=True)     # no allocation inside the timed loop
        run = lambda: cr.propagate(G, E0u, E0i, LAYERS, buffers=bufs)
        nnz_local = G.nnz
    else:
        from coldrec_b200.dist import RowPartitionedGraph
        PG = RowPartitionedGraph(G.rowptr.cpu().numpy(), G.col.cpu().numpy(), G.val.cpu().numpy(), device,
                                 segments=(n_users, n_items))
        PG.local.plan(D)
        E0 = torch.cat([E0u, E0i])
        if args.peer_store == "st":
            os.environ["CR_SPMM_PEER_ST"] = "1"
        exchange = ("SpMM epilogue -> shared-memory staging -> TMA bulk stores (cp.async.bulk) into the readers' tables over NVLink "
                    "(fused all-gather)" if args.peer_store == "tma" and not args.multicast else
                    "SpMM epilogue 16-byte peer stores over NVLink (fused all-gather)")
        try:
            PG.enable_p2p(D)
            rep = None if args.prop_result == "full" else (0,)
            # copy=False: a view of the peer-mapped result table, like the fresh tensors of N=1
            run = lambda: PG.propagate_p2p(E0, LAYERS, copy=False, sparse=not args.dense_exchange, replicate_result=rep,
                                           multicast=args.multicast)
            run()
            exchange += (", dense" if args.dense_exchange else f", rows sent only to the GPUs that gather them ({PG.need_copies:.2f} of {world} "
                         f"copies per row)") + f", result: {args.prop_result}" + (
                             ", rows wanted everywhere via NVLS multimem.st" if (PG.has_multicast and args.multicast) else ", unicast only")
        except Exception as ex:      # symmetric memory unavailable on this box: NCCL all-gather after each layer
            print(f"[bench] peer-store path failed on rank {rank}: {type(ex).__name__}: {ex}", file=sys.stderr, flush=True)
            exchange = f"NCCL all-gather per layer (peer path unavailable: {type(ex).__name__}: {str(ex)[:80]})"
            run = lambda: PG.propagate(E0, LAYERS)
        nnz_local = PG.local.nnz
    for _ in range(W):
        run()
    barrier(world, device)
    launches0 = lib.cr_launch_count()
    lib.cr_profile_enable(1)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(device.index) as clk:
        ev0.record()
        for _ in range(Ksteps):
            res = run()
        ev1.record()
        barrier(world, device)
        ms = max_over_ranks(ev0.elapsed_time(ev1), device, world)
        tot, cnt = ctypes.c_double(), ctypes.c_int()
        lib.cr_profile_read(1, ctypes.byref(tot), ctypes.byref(cnt))
        lib.cr_profile_enable(0)
        launches = lib.cr_launch_count() - launches0
        if world == 1:
            clk.top_up(run)
        elif ms < 600:      # collective step: every rank runs the same number of extra (untimed) steps, ~0.6 s of the same load
            n0 = len(clk.lines)
            for _ in range(min(200, int(600 * Ksteps / max(ms, 1e-3)) + 1)):
                run()
            torch.cuda.synchronize()
            time.sleep(0.05)
            clk.tail = len(clk.lines) - n0
    nnz = G.nnz
    # nnz*(4 idx + 4 val + 4d) + N*4d (write E_k+1) + 2*N*4d (layer-mean RMW) + 8(N+1) rowptr, per layer
    bytes_layer = nnz * (8 + 4 * D) + 3 * N * 4 * D + 8 * (N + 1)
    step_gbs = bytes_layer * LAYERS / (ms / Ksteps * 1e-3) / 1e9 / world
    rows_ms = tot.value / max(cnt.value, 1)
    out = dict(metric="LightGCN edges/sec (stored nonzeros x layers / s)", value=round(nnz * LAYERS * Ksteps / (ms * 1e-3), 1),
               unit="edges/s", n_gpus=world, steps=Ksteps, warmup=W, ms_per_step=round(ms / Ksteps, 3), higher_is_better=True,
               scaling="strong", dtype="f32", data="synthetic",
               config={"workload": f"C4 LightGCN {LAYERS}-layer propagation: {n_users} users + {n_items} items, {args.graph_edges} "
                                   f"interactions (nnz(A)={nnz}), d={D}, fused layer mean",
                       "parallelism": f"row-partitioned x{world}, {exchange}" if world > 1 else "single GPU",
                       "l2": "embedding table %.1f GB >> L2; no flush" % (N * D * 4 / 2**30)},
               gpu_launches=int(launches),
               roofline={"bound": "hbm", "kernel": "spmm_rows_grouped_kernel (rows + long-row chunks in one launch; + spmm_long_reduce_kernel)", "achieved": round(step_gbs, 1),
                         "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": round(step_gbs / pk["hbm_gbs"], 4),
                         "traffic": None,
                         "peak_source": f"{pk_src} hbm_gbs (copy bandwidth)", "bytes_per_layer": bytes_layer,
                         "rows_kernel_ms": round(rows_ms, 3), "rows_kernel_launches": cnt.value,
                         "share_of_step": round(tot.value / ms, 4)},
               clocks=clk.summary())
    if world == 1 and args.graph_edges == GRAPH_EDGES:
        out["roofline"]["traffic"], note = ncu_traffic("spmm_rows_grouped_kernel", "spmm.cu")
        if note:
            out["roofline"]["traffic_note"] = note
    if world > 1:
        out["check"] = check_partitioned_propagation(G, E0u, E0i, res, PG, args, device, world)
        out["nvlink"] = nvlink_bytes(PG, args, ms / Ksteps, device, world)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_spmm_baseline(G, E0u, E0i)
        out["gpu_library_baseline"] = library_spmm_baseline(G, E0u, E0i)
        torch.cuda.empty_cache()
    if world == 1 and not args.no_train:
        out["train_step"] = run_train_step(args, device, G, E0u, E0i, pk, lib)
    return out


def run_train_step(args, device, G, E0u, E0i, pk, lib):
    """One LightGCN optimisation step (model/LightGCN.py:21-28) on the same graph: device sampler -> forward propagation ->
    fused BPR loss + row gradients -> backward propagation (same SpMM, symmetric adjacency) -> Adam.  Reported next to the
    propagation line; the sampled batch changes every step."""
    import coldrec_b200 as cr
    n_users, n_items = E0u.shape[0], E0i.shape[0]
    N, nnz = n_users + n_items, G.nnz
    # training pairs = the user->item half of the adjacency (one per stored interaction)
    rp_u = G.rowptr[:n_users + 1]
    pu = torch.repeat_interleave(torch.arange(n_users, device=device, dtype=torch.int32), (rp_u[1:] - rp_u[:-1]))
    pi = (G.col[:int(rp_u[-1])] - n_users).to(torch.int32)
    smp = cr.PairwiseSampler(pu, pi, n_users, n_items, seed=7)
    del pu, pi
    step = cr.BprTrainStep(G, E0u, E0i, LAYERS, 1e-3, 1e-4)
    B, W, Ksteps = args.train_batch, max(args.warmup, 3), max(args.steps, 3)
    buf = torch.empty((3, B), dtype=torch.int32, device=device)
    def one(k):
        u, i, j = smp.batch(0, (k * B) % max(smp.n_pairs - B, 1), B, out=buf)
        return step.step(u, i, j)
    for k in range(W):
        one(k)
    torch.cuda.synchronize(device)
    l0 = lib.cr_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(W, W + Ksteps):
        loss = one(k)
    e1.record()
    torch.cuda.synchronize(device)
    ms = e0.elapsed_time(e1) / Ksteps
    # algorithmic bytes: two propagations (SURVEY 8d per layer) + Adam (28 B/element) + gradient-table clear (4 B/element)
    bytes_layer = nnz * (8 + 4 * D) + 3 * N * 4 * D + 8 * (N + 1)
    step_bytes = 2 * LAYERS * bytes_layer + N * D * 32
    return {"metric": "LightGCN training steps/sec (sample + forward + BPR + backward + Adam)", "value": round(1e3 / ms, 3), "unit": "steps/s",
            "ms_per_step": round(ms, 3), "batch": B, "edges_per_s": round(2 * LAYERS * nnz / (ms * 1e-3), 1),
            "gpu_launches": int(lib.cr_launch_count() - l0), "steps": Ksteps, "warmup": W,
            "roofline": {"bound": "hbm", "achieved": round(step_bytes / (ms * 1e-3) / 1e9, 1), "peak": pk["hbm_gbs"], "unit": "GB/s",
                         "frac": round(step_bytes / (ms * 1e-3) / 1e9 / pk["hbm_gbs"], 4), "bytes_per_step": step_bytes},
            "loss": [round(x, 6) for x in loss.cpu().tolist()[:3]], "sampler_exhausted": int(smp.n_exhausted.item())}




# ------------------------------------------------------------------------------------------- robustness lines (N = 1)
def run_robustness(args, lib, user_tab, item_tab, plans, device, pk, steps=3, warmup=1):
    """The same sweep on inputs the headline line does not cover: the 'warm' / 'cold' settings with 20 % cold
    items through BOTH mask paths (item flags tested inside the kernel's bitmap producer; flagged items compacted away
    before the sweep, ids recovered through ``item_gids``), 1 % exactly duplicated item rows (score ties), log-normal item
    norms (sigma = 1: stresses the margin proof, whose eps scales with the largest item norm) and the same table sorted by
    ascending norm (the thresholds keep rising to the end of the sweep).  ``n_refined`` = queries whose margin proof failed
    and were re-ranked by the exact fp32 kernel, summed over ALL timed steps."""
    import ctypes
    from coldrec_b200 import ops
    from coldrec_b200.scoring import EvalPlan, FullRankScorer, FLAG_COLD, FLAG_WARM
    n_items, n_q = item_tab.shape[0], plans[0].n_q
    g = torch.Generator(device=device).manual_seed(77)
    flags = torch.where(torch.rand(n_items, device=device, generator=g) < 0.2, FLAG_COLD, FLAG_WARM).to(torch.uint8)
    tf32_peak = pk["bf16_tflops_sustained"] / 2.0
    lines = []

    plans = plans[:2]            # two eval batches, both seen in the warm-up: per-plan host state (memoised remaps) is steady state

    def timed(name, fn, n_swept):
        for k in range(max(warmup, len(plans))):
            fn(plans[k % len(plans)])
        torch.cuda.synchronize(device)
        lib.cr_profile_enable(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        nref = torch.zeros(1, dtype=torch.int64, device=device)
        e0.record()
        for k in range(steps):
            for r in fn(plans[(warmup + k) % len(plans)]):
                nref += r.to(torch.int64)
        e1.record()
        torch.cuda.synchronize(device)
        tot, cnt = ctypes.c_double(), ctypes.c_int()
        lib.cr_profile_read(0, ctypes.byref(tot), ctypes.byref(cnt)); lib.cr_profile_enable(0)
        ms = e0.elapsed_time(e1) / steps
        sweep_ms = tot.value / max(cnt.value, 1)
        tfl = 2.0 * n_q * n_swept * D / (sweep_ms * 1e-3) / 1e12 if sweep_ms > 0 else 0.0
        lines.append({"case": name, "users_per_s": round(n_q / (ms * 1e-3), 1), "ms_per_step": round(ms, 3), "items_swept": int(n_swept),
                      "sweep_ms": round(sweep_ms, 3), "sweep_tflops": round(tfl, 1), "sweep_frac_of_tf32_sustained": round(tfl / tf32_peak, 4),
                      "n_refined_all_steps": int(nref.item()), "queries_all_steps": n_q * steps})

    def kernel_flags(excl):
        def fn(p):
            s, i, nref = ops.score_topk(user_tab, item_tab, K, user_ids=p.user_ids, mask_rowptr=p.mask_rowptr, mask_col=p.mask_col,
                                        item_flags=flags if excl else None, flag_exclude=excl, precision=ops.SCORE_TF32_CHECKED)
            ops.rank_metrics(i, p.gt_rowptr, p.gt_col, TOPN)
            return [nref]
        return fn

    def compacted(excl):
        fr = FullRankScorer(K, ops.SCORE_TF32_CHECKED)
        # one plan object per eval batch and setting, as the trainer memoises them (_get_eval_cache): the scorer keeps the
        # mask-to-row-number remap of a compacted table per (plan, selection)
        qplans = {id(p): EvalPlan(None, p.user_ids, p.mask_rowptr, p.mask_col, p.gt_rowptr, p.gt_col, excl) for p in plans}
        def fn(p):
            q = qplans[id(p)]
            s, i = fr.topk([(user_tab, item_tab, None)], q, flags)
            ops.rank_metrics(i, p.gt_rowptr, p.gt_col, TOPN)
            return fr.n_refined
        return fn
    n_warm, n_cold = int((flags == FLAG_WARM).sum()), int((flags == FLAG_COLD).sum())
    timed("all (no flag mask; the headline case)", kernel_flags(0), n_items)
    timed("warm setting, 20% cold items masked inside the kernel (flag byte per item)", kernel_flags(FLAG_COLD), n_items)
    timed("warm setting, cold items compacted away before the sweep (item_gids)", compacted(FLAG_COLD), n_warm)
    timed("cold setting, 80% warm items masked inside the kernel", kernel_flags(FLAG_WARM), n_items)
    timed("cold setting, warm items compacted away before the sweep (item_gids)", compacted(FLAG_WARM), n_cold)
    # the exact fp32 FFMA kernel (CR_SCORE_EXACT_F32: the product path for d not in {64, 128}, K > 52, and every query whose
    # margin proof fails): 4,096 of the step's queries against the whole catalogue; bound = fp32 FMA rate of the CUDA cores
    # (148 SMs x 128 lanes x 2 x SM clock: ~72 TFLOP/s at 1.9 GHz), not the tensor pipe
    try:
        nx = min(4096, n_q)
        px = plans[0]
        sub = px.slice(0, nx)
        fx = lambda: ops.score_topk(user_tab, item_tab, K, user_ids=sub.user_ids, mask_rowptr=sub.mask_rowptr, mask_col=sub.mask_col,
                                    precision=ops.SCORE_EXACT_F32)
        ex_ms = _event_ms(fx, iters=2, warm=1)
        ex_tfl = 2.0 * nx * n_items * D / (ex_ms * 1e-3) / 1e12
        lines.append({"case": f"exact fp32 FFMA kernel (score_topk_exact_kernel), {nx} queries x {n_items} items", "users_per_s": round(nx / (ex_ms * 1e-3), 1),
                      "ms_per_step": round(ex_ms, 3), "items_swept": n_items, "sweep_tflops": round(ex_tfl, 1), "bound": "fp32 FFMA",
                      "fp32_ffma_peak_tflops": 72.0, "frac_of_fp32_ffma_peak": round(ex_tfl / 72.0, 3)})
    except Exception as ex:
        lines.append({"case": "exact fp32 FFMA kernel", "error": f"{type(ex).__name__}: {str(ex)[:120]}"})
    saved = item_tab.clone()
    try:
        nd = n_items // 100
        perm = torch.randperm(n_items, device=device, generator=g)
        item_tab[perm[:nd]] = item_tab[perm[nd:2 * nd]]
        timed("1% of the item rows exact duplicates of other rows (score ties)", kernel_flags(0), n_items)
        item_tab.copy_(saved)
        item_tab *= torch.exp(torch.randn(n_items, 1, device=device, generator=g))
        timed("log-normal item norms (sigma = 1)", kernel_flags(0), n_items)
        order = torch.argsort((item_tab * item_tab).sum(1))
        item_tab.copy_(item_tab[order])
        del order
        timed("log-normal item norms, table sorted by ascending norm", kernel_flags(0), n_items)
    finally:
        item_tab.copy_(saved)
        del saved
    return lines


# ------------------------------------------------------------------------------------------- d = 128 line (N = 1)
def run_d128(args, lib, plans, device, pk, steps=3, warmup=2):
    """VBPR / AMR score with two inner products, P.Q^T + P2.Q2^T = [P|P2].[Q|Q2]^T (model/VBPR.py:68-75): one sweep at d = 128
    through the d = 128 instantiation of the tcgen05 kernel (64-item tiles, 2 x 16 MMAs per tile), same users, masks and
    catalogue size as the headline line."""
    import ctypes
    from coldrec_b200 import ops
    g = torch.Generator(device=device).manual_seed(128)
    U = torch.randn(args.n_users, 128, device=device, generator=g) * 0.09
    I = torch.randn(args.n_items, 128, device=device, generator=g) * 0.09
    n_q = plans[0].n_q

    def step(p):
        s, i, nref = ops.score_topk(U, I, K, user_ids=p.user_ids, mask_rowptr=p.mask_rowptr, mask_col=p.mask_col, precision=ops.SCORE_TF32_CHECKED)
        ops.rank_metrics(i, p.gt_rowptr, p.gt_col, TOPN)
        return nref
    for k in range(warmup):
        step(plans[k % len(plans)])
    torch.cuda.synchronize(device)
    lib.cr_profile_enable(1)
    l0 = lib.cr_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    nref = torch.zeros(1, dtype=torch.int64, device=device)
    e0.record()
    for k in range(steps):
        nref += step(plans[(warmup + k) % len(plans)]).to(torch.int64)
    e1.record()
    torch.cuda.synchronize(device)
    tot, cnt = ctypes.c_double(), ctypes.c_int()
    lib.cr_profile_read(0, ctypes.byref(tot), ctypes.byref(cnt)); lib.cr_profile_enable(0)
    ms, sweep_ms = e0.elapsed_time(e1) / steps, tot.value / max(cnt.value, 1)
    tf32_peak = pk["bf16_tflops_sustained"] / 2.0
    tfl = 2.0 * n_q * args.n_items * 128 / (sweep_ms * 1e-3) / 1e12 if sweep_ms > 0 else 0.0
    del U, I
    torch.cuda.empty_cache()
    return {"metric": "full-rank users/sec (top-20 over catalog), d = 128", "value": round(n_q / (ms * 1e-3), 1), "unit": "users/s",
            "ms_per_step": round(ms, 3), "steps": steps, "warmup": warmup, "gpu_launches": int(lib.cr_launch_count() - l0),
            "config": {"workload": f"VBPR/AMR-shaped two-product scoring as one d=128 sweep: {args.n_users} users x {args.n_items} items, "
                                   f"{n_q} users/step, same masks / K as the headline line"},
            "roofline": {"bound": "tensor", "kernel": "score_sweep_tc_kernel<128>", "achieved": round(tfl, 1), "peak": round(tf32_peak, 1),
                         "unit": "TFLOP/s", "frac": round(tfl / tf32_peak, 4), "launch_ms": round(sweep_ms, 3), "launches": cnt.value,
                         "flop_per_launch": 2.0 * n_q * args.n_items * 128},
            "n_refined_all_steps": int(nref.item()), "queries_all_steps": n_q * steps}

# ------------------------------------------------------------------------------------------- multi-GPU result checks
def check_sharded_ids(args, scorer, user_tab, item_shard, ib, plan, S, device, world, n_sample=4096):
    """N>1: the merged lists this rank holds for the first `n_sample` users of its slice must equal, bit for bit, ONE
    un-sharded sweep of the same users over the whole catalogue (rebuilt here from the shard seeds) — ids and scores."""
    import torch.distributed as dist
    from coldrec_b200 import ops
    from coldrec_b200.dist import shard_range
    s, i = scorer.topk(user_tab, item_shard, ib, plan)
    lo, hi = scorer.user_slice(plan.n_q)
    n = min(n_sample, hi - lo)
    parts = []
    for sh in range(S):      # the whole catalogue, shard by shard, from the generators the ranks used
        b, e = shard_range(args.n_items, sh, S)
        parts.append(torch.randn(e - b, D, device=device, generator=torch.Generator(device=device).manual_seed(1000 + sh)) * 0.125)
    full = torch.cat(parts) if len(parts) > 1 else parts[0]
    del parts
    sub = plan.slice(lo, lo + n)
    rs, ri, _ = ops.score_topk(user_tab, full, K, user_ids=sub.user_ids, mask_rowptr=sub.mask_rowptr, mask_col=sub.mask_col,
                               precision=ops.SCORE_TF32_CHECKED)
    ok = torch.tensor([int(torch.equal(ri, i[:n])), int(torch.equal(rs, s[:n]))], device=device)
    del full
    torch.cuda.empty_cache()
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    return {"ids_equal": bool(ok[0].item()), "scores_equal": bool(ok[1].item()), "ids_checked_users_per_rank": n,
            "ids_checked_against": f"one un-sharded sweep over all {args.n_items} items on every rank"}


def check_partitioned_propagation(G, E0u, E0i, res, PG, args, device, world):
    """N>1: every row this rank holds of the partitioned result vs a single-GPU propagation of the same graph run on
    this rank (every rank has the whole graph): norm-wise error max|d| / max|ref|, max over ranks."""
    import coldrec_b200 as cr
    import torch.distributed as dist
    n_users = E0u.shape[0]
    ru, ri = cr.propagate(G, E0u, E0i, LAYERS)
    scale = max(ru.abs().max().item(), ri.abs().max().item())
    if args.prop_result == "full":
        err = max((res[:n_users] - ru).abs().max().item(), (res[n_users:] - ri).abs().max().item())
        rows = res.shape[0]
    else:                    # user rows everywhere, item rows with their owner
        (ub, ue), (ib_, ie_) = PG.parts[PG.rank]
        err = max((res[:n_users] - ru).abs().max().item(), (res[ib_:ie_] - ri[ib_ - n_users:ie_ - n_users]).abs().max().item())
        rows = n_users + (ie_ - ib_)
    deg = G.rowptr[1:] - G.rowptr[:-1]
    t_ = torch.tensor([err / scale], dtype=torch.float64, device=device)
    dist.all_reduce(t_, op=dist.ReduceOp.MAX)
    del ru, ri
    torch.cuda.empty_cache()
    return {"prop_rel_err": float(t_.item()), "rows_checked_per_rank": int(rows), "longest_row_nnz": int(deg.max().item()),
            "against": "single-GPU cr.propagate of the same graph on every rank (all rows, incl. the long-row split path)",
            "tolerance": 1e-5}


def nvlink_bytes(PG, args, ms_step, device, world):
    """Bytes this step's fused all-gather moves over NVLink, from the need masks (exact: one d*4-byte row per remote reader
    and layer), and the per-GPU egress rate they imply; 770 GB/s per direction is the measured peer-copy rate
    (B200_PROFILING.md)."""
    import torch.distributed as dist
    self_bit = 1 << PG.rank
    def remote_copies(need):
        if need is None:
            return PG.n_local * (world - 1)
        m = need[:PG.n_local].to(torch.int32) & ~self_bit
        return int(sum(((m >> p) & 1).sum().item() for p in range(world)))
    inner = remote_copies(None if args.dense_exchange else PG._need)
    if args.prop_result == "full":
        last = PG.n_local * (world - 1)
    else:
        last = remote_copies(PG._need_last.get((0,)))
    egress = (inner * (LAYERS - 1) + last) * D * 4
    t_ = torch.tensor([egress], dtype=torch.float64, device=device)
    tmax = t_.clone()
    dist.all_reduce(t_, op=dist.ReduceOp.SUM)
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    return {"bytes_per_step_all_gpus": int(t_.item()), "egress_bytes_per_step_max_gpu": int(tmax.item()),
            "egress_gbs_max_gpu": round(tmax.item() / (ms_step * 1e-3) / 1e9, 1), "peer_copy_peak_gbs": 770.0,
            "nvlink_bound_ms": round(tmax.item() / 770e9 * 1e3, 3)}

# ------------------------------------------------------------------------------------------- the reference's GPU path
def _event_ms(fn, iters=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def library_score_baseline(user_tab, item_tab, plan_d, batch=512):
    """What ColdRec itself runs with --use_gpu true ("existing Blackwell library path"): torch.matmul with TF32
    off (model/MF.py:62), the per-user index_put mask loop (model/BaseRecommender.py:175-177) and torch.topk (:182), on one
    batch of this step's users.  Reported next to the CPU baseline, never part of the product path."""
    try:
        users = plan_d["user_ids"][:batch].long()
        rp = plan_d["mask_rowptr"][:batch + 1].tolist()
        rated = [plan_d["mask_col"][rp[j]:rp[j + 1]].long() for j in range(batch)]

        def one_batch():
            cand = torch.matmul(user_tab[users], item_tab.transpose(0, 1))
            for j in range(batch):
                cand[j, rated[j]] = -10e8
            return torch.topk(cand, K, dim=1, largest=True, sorted=True)
        ms = _event_ms(one_batch)
        return {"value": round(batch / ms * 1e3, 1), "unit": "users/s", "kind": "reference torch ops on this GPU",
                "sample": f"{batch}-user batches x {item_tab.shape[0]} items (matmul fp32 + index_put loop + topk), {ms:.1f} ms per batch"}
    except Exception as ex:      # e.g. out of memory on a smaller GPU: the bench line must still come out
        torch.cuda.empty_cache()
        return {"error": f"{type(ex).__name__}: {str(ex)[:120]}"}


def library_spmm_baseline(G, E0u, E0i):
    """torch.sparse.mm on the coalesced int64 COO tensor (util/databuilder.py:959-962, model/LightGCN.py:86-96) + stack + mean."""
    try:
        N = G.n_rows
        rows = torch.repeat_interleave(torch.arange(N, device=G.col.device), G.rowptr[1:] - G.rowptr[:-1])
        A = torch.sparse_coo_tensor(torch.stack([rows, G.col.long()]), G.val, (N, N)).coalesce()
        del rows
        E0 = torch.cat([E0u, E0i])

        def forward():
            ego, layers = E0, [E0]
            for _ in range(LAYERS):
                ego = torch.sparse.mm(A, ego)
                layers.append(ego)
            return torch.mean(torch.stack(layers, dim=1), dim=1)
        ms = _event_ms(forward)
        return {"value": round(G.nnz * LAYERS / ms * 1e3, 1), "unit": "edges/s", "kind": "reference torch ops on this GPU",
                "sample": f"torch.sparse.mm (COO int64) x {LAYERS} + stack + mean, {ms:.1f} ms per propagation"}
    except Exception as ex:
        torch.cuda.empty_cache()
        return {"error": f"{type(ex).__name__}: {str(ex)[:120]}"}


# ------------------------------------------------------------------------------------------- CPU arms (oracle port)
def cpu_score_baseline(user_tab, item_tab, plan_d, n_sample, args):
    """The reference's CPU path on a bounded sample: its own MF.batch_predict + _evaluate + ranking_evaluation (baseline/_ref,
    kind "reference"), or the oracle restatement when the reference tree did not travel (kind "port")."""
    import copy
    torch.set_num_threads(host_threads())
    U, I = user_tab.cpu(), item_tab.cpu()
    a = copy.copy(args)
    a.cpu_sample_users, a.n_items, a.n_users = n_sample, I.shape[0], U.shape[0]
    step, kind, sample = reference_step_factory(a, U, I)
    rp = plan_d["mask_rowptr"][:n_sample + 1].cpu()
    grp = plan_d["gt_rowptr"][:n_sample + 1].cpu()
    p = dict(user_ids=plan_d["user_ids"][:n_sample].cpu(), mask_rowptr=rp, mask_col=plan_d["mask_col"][:int(rp[-1])].cpu(),
             gt_rowptr=grp, gt_col=plan_d["gt_col"][:int(grp[-1])].cpu())
    t0 = time.perf_counter()
    step(p)
    dt = time.perf_counter() - t0
    return {"value": round(n_sample / dt, 2), "unit": "users/s", "cores": torch.get_num_threads(), "kind": kind,
            "sample": f"{sample}, {dt:.1f} s"}


def cpu_spmm_baseline(G, E0u, E0i, frac=0.05):
    """torch.sparse.mm (COO, int64 indices) on the host cores, on a contiguous row sample of the adjacency."""
    torch.set_num_threads(host_threads())
    n_rows = int(G.n_rows * frac)
    lo, hi = 0, int(G.rowptr[n_rows])
    rows = torch.repeat_interleave(torch.arange(n_rows, device=G.rowptr.device), (G.rowptr[1:n_rows + 1] - G.rowptr[:n_rows]))
    idx = torch.stack([rows, G.col[lo:hi].long()]).cpu()
    A = torch.sparse_coo_tensor(idx, G.val[lo:hi].cpu(), (n_rows, G.n_cols)).coalesce()
    X = torch.cat([E0u, E0i]).cpu()
    t0 = time.perf_counter()
    torch.sparse.mm(A, X)
    dt = time.perf_counter() - t0
    return {"value": round((hi - lo) / dt, 1), "unit": "edges/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"one layer over the first {n_rows} rows ({hi - lo} nonzeros, {frac:.0%} of rows) of the C4 adjacency, {dt:.1f} s"}


class _CatalogueData:
    """What ``BaseColdStartTrainer._evaluate`` / ``_get_eval_cache`` / ``MF.batch_predict`` touch of a data builder
    (model/BaseRecommender.py:109-188, model/MF.py:58-63), for a catalogue too large for the reference's dict builders:
    raw ids ARE dense ids, train items only for the sampled users."""

    def __init__(self, n_users, n_items, plan):
        from collections import defaultdict
        self.user_num, self.item_num = n_users, n_items
        self.id2item = range(n_items)
        self.training_set_u = defaultdict(dict)
        uids, rp, col = plan["user_ids"].tolist(), plan["mask_rowptr"].tolist(), plan["mask_col"].tolist()
        grp, gcol = plan["gt_rowptr"].tolist(), plan["gt_col"].tolist()
        self.overall_test_set = {}
        for j, u in enumerate(uids):
            self.training_set_u[u] = dict.fromkeys(col[rp[j]:rp[j + 1]], 1.0)
            self.overall_test_set[u] = dict.fromkeys(gcol[grp[j]:grp[j + 1]], 1.0)
        self.mapped_cold_item_idx = self.mapped_warm_item_idx = []

    def get_user_id_list(self, users):
        return np.asarray(users)

    def get_item_id_list(self, items):
        return np.asarray(items)


def reference_step_factory(args, U, I):
    """One step of the reference arm.  With baseline/_ref present (it travels with the snapshot): the reference's OWN
    ``MF.batch_predict`` + ``BaseColdStartTrainer._evaluate`` + ``util.evaluator.ranking_evaluation``, unmodified, with the
    eval batch (``--bs``) = the sampled users: a (sample x 10M) fp32 score matrix, 5 GB at 128 users — the literal code
    path, which is why the sample is small.  Otherwise the oracle port (item-chunked restatement)."""
    import types
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    try:
        import refimport
    finally:
        sys.path.pop(0)
    n_q = args.cpu_sample_users
    if refimport.available():
        refimport.import_reference()
        from model.BaseRecommender import BaseColdStartTrainer
        from model.MF import MF
        from util.evaluator import ranking_evaluation

        def step(p):
            data = _CatalogueData(args.n_users, args.n_items, p)
            ns = types.SimpleNamespace(topN="10,20", model="MF", dataset="syn", emb_size=D, epochs=0, bs=n_q, lr=1e-3, reg=1e-4,
                                       early_stop=0, eval_every=1, cold_object="item", save_emb=False)
            tr = object.__new__(MF)          # MF.__init__ would allocate (and xavier-initialise) a second 10M x 64 table
            BaseColdStartTrainer.__init__(tr, types.SimpleNamespace(args=ns, data=data, device=torch.device("cpu")))
            tr.user_emb, tr.item_emb = U, I
            rec = tr.test("all")
            measure, perf = ranking_evaluation(data.overall_test_set, rec, TOPN)
            return perf
        return step, "reference", (f"{n_q} users/step x {args.n_items} items: the reference's own MF.batch_predict + _evaluate (one "
                                   f"{n_q}-user batch) + ranking_evaluation from baseline/_ref")
    from oracle import coldrec_oracle as O

    def step(p):
        s, i = O.evaluate_topk_dense_chunked(U, I, p["user_ids"].numpy(), p["mask_rowptr"].numpy(), p["mask_col"].numpy().astype(np.int64),
                                             None, K, user_batch=min(128, n_q), item_chunk=1 << 20)
        return O.metrics_from_topk(i, p["gt_rowptr"].numpy(), p["gt_col"].numpy().astype(np.int64), TOPN)
    return step, "port", f"{n_q} users/step x {args.n_items} items, item-chunked torch CPU matmul + mask + topk + metrics (oracle port)"


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the host cores, all host threads, a bounded
    sample per step (see reference_step_factory)."""
    rank, _, world = dist_env()
    if rank != 0:
        return
    torch.set_num_threads(host_threads())
    n_q = args.cpu_sample_users
    g = torch.Generator().manual_seed(1)
    U = torch.randn(args.n_users, D, generator=g) * 0.125
    I = torch.randn(args.n_items, D, generator=g) * 0.125
    plans = make_step_plans(args.warmup + args.steps, n_q, args.n_users, args.n_items, 6, torch.device("cpu"))
    step, kind, sample = reference_step_factory(args, U, I)
    for p in plans[:args.warmup]:
        step(p)
    t0 = time.perf_counter()
    for p in plans[args.warmup:]:
        perf = step(p)
    dt = time.perf_counter() - t0
    value = n_q * args.steps / dt
    cores = torch.get_num_threads()
    line = dict(impl="reference", metric="full-rank users/sec (top-20 over catalog)", value=round(value, 2), unit="users/s",
                n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=round(dt / args.steps * 1e3, 1),
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                config={"workload": score_workload(args, args.users_per_step * world), "parallelism": "host cores (CPU reference path)",
                        "users_per_step": args.users_per_step * world, "n_items": args.n_items, "K": K,
                        "sampled_users_per_step": n_q},
                cpu_baseline={"value": round(value, 2), "unit": "users/s", "cores": cores, "kind": kind, "sample": sample},
                e2e={"value": round(value, 2), "unit": "users/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                gpu_launches=0, check={"ndcg@20": perf[1][3]})
    if not args.no_configs and world == 1:
        import bench_configs
        line["extra"] = {}
        for key in [k for k in args.configs.split(",") if k in bench_configs.CONFIGS]:
            try:
                line["extra"][key] = bench_configs.reference_config_line(key, cores)
            except Exception as ex:
                line["extra"][key] = {"error": f"{type(ex).__name__}: {str(ex)[:200]}"}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
