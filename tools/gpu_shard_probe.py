"""Probe: the tcgen05 sweep at the per-GPU shapes of the item-sharded runs (one GPU emulating rank `r` of W), for the
library named by CR_LIB_PATH.  Prints TFLOP/s of the sweep kernel (library CUDA events) per shape."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from coldrec_b200 import ops, _lib

dev = torch.device("cuda:0")
lib = _lib.load()
N_USERS, N_ITEMS, MASK = 1_000_000, 10_000_000, 100
g = torch.Generator(device=dev).manual_seed(1)
U = torch.randn(N_USERS, 64, device=dev, generator=g) * 0.125
shapes = [(1, 75_776), (8, 606_208), (4, 303_104), (2, 151_552)]
if len(sys.argv) > 1:
    shapes = [s for s in shapes if str(s[0]) in sys.argv[1].split(",")]
for W, n_q in shapes:
    n_loc = N_ITEMS // W
    base = (W // 2) * n_loc if W > 1 else 0
    I = torch.randn(n_loc, 64, device=dev, generator=g) * 0.125
    uids = (torch.arange(n_q, device=dev, dtype=torch.int64) % N_USERS).to(torch.int32)
    mrp = torch.arange(0, (n_q + 1) * MASK, MASK, device=dev, dtype=torch.int64)
    x = torch.sort(torch.randint(0, N_ITEMS - MASK, (n_q, MASK), device=dev, generator=g), dim=1).values
    mc = (x + torch.arange(MASK, device=dev)).to(torch.int32).flatten().contiguous()
    del x
    run = lambda: ops.score_topk(U, I, 20, user_ids=uids, item_id_base=base, mask_rowptr=mrp, mask_col=mc, precision=ops.SCORE_TF32_CHECKED)
    for _ in range(2):
        s, i, nref = run()
    torch.cuda.synchronize()
    lib.cr_profile_enable(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 4
    e0.record()
    for _ in range(iters):
        s, i, nref = run()
    e1.record(); torch.cuda.synchronize()
    tot, cnt = ctypes.c_double(), ctypes.c_int()
    lib.cr_profile_read(0, ctypes.byref(tot), ctypes.byref(cnt)); lib.cr_profile_enable(0)
    sweep_ms = tot.value / max(cnt.value, 1)
    flops = 2.0 * n_q * n_loc * 64
    print(json.dumps(dict(lib=os.path.basename(os.environ.get("CR_LIB_PATH", "default")), W=W, n_q=n_q, n_items=n_loc,
                          sweep_ms=round(sweep_ms, 3), call_ms=round(e0.elapsed_time(e1) / iters, 3), tflops=round(flops / sweep_ms / 1e9, 1),
                          n_refined=int(nref.item()), id_sum=int(i.long().sum().item()))), flush=True)
    del I, mc, s, i
    torch.cuda.empty_cache()
