"""Probe: the sweep with the warm / cold flag masks tested inside the kernel (20 % cold items), at the 10M-item and the
2.5M-item shard shapes, for the library named by CR_LIB_PATH.  Prints TFLOP/s of the sweep kernel per setting."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from coldrec_b200 import ops, _lib

dev = torch.device("cuda:0")
lib = _lib.load()
N_USERS, N_ITEMS, MASK = 1_000_000, 10_000_000, 100
g = torch.Generator(device=dev).manual_seed(1)
U = torch.randn(N_USERS, 64, device=dev, generator=g) * 0.125
flags_all = torch.where(torch.rand(N_ITEMS, device=dev, generator=g) < 0.2, 1, 2).to(torch.uint8)
for W, n_q in ((1, 75_776), (4, 303_104)):
    n_loc = N_ITEMS // W
    base = (W // 2) * n_loc if W > 1 else 0
    I = torch.randn(n_loc, 64, device=dev, generator=g) * 0.125
    uids = (torch.arange(n_q, device=dev, dtype=torch.int64) % N_USERS).to(torch.int32)
    mrp = torch.arange(0, (n_q + 1) * MASK, MASK, device=dev, dtype=torch.int64)
    x = torch.sort(torch.randint(0, N_ITEMS - MASK, (n_q, MASK), device=dev, generator=g), dim=1).values
    mc = (x + torch.arange(MASK, device=dev)).to(torch.int32).flatten().contiguous()
    del x
    for name, excl in (("all", 0), ("warm", 1), ("cold", 2)):
        run = lambda: ops.score_topk(U, I, 20, user_ids=uids, item_id_base=base, mask_rowptr=mrp, mask_col=mc,
                                     item_flags=flags_all if excl else None, flag_exclude=excl, precision=ops.SCORE_TF32_CHECKED)
        for _ in range(2):
            s, i, nref = run()
        torch.cuda.synchronize()
        lib.cr_profile_enable(1)
        for _ in range(3):
            s, i, nref = run()
        torch.cuda.synchronize()
        tot, cnt = ctypes.c_double(), ctypes.c_int()
        lib.cr_profile_read(0, ctypes.byref(tot), ctypes.byref(cnt)); lib.cr_profile_enable(0)
        sweep_ms = tot.value / max(cnt.value, 1)
        print(json.dumps(dict(lib=os.path.basename(os.environ.get("CR_LIB_PATH", "default")), W=W, setting=name, n_items=n_loc,
                              sweep_ms=round(sweep_ms, 3), tflops=round(2.0 * n_q * n_loc * 64 / sweep_ms / 1e9, 1),
                              n_refined=int(nref.item()), id_sum=int(i.long().sum().item()))), flush=True)
    del I, mc
    torch.cuda.empty_cache()
