"""GPU probe: per-unit timeline of the tcgen05 sweep (CR_TC_DEBUG_MODE=8)."""
import ctypes, os, sys
os.environ["CR_TC_DEBUG_MODE"] = os.environ.get("CR_TC_DEBUG_MODE", "8")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from coldrec_b200 import ops, _lib

dev = torch.device("cuda:0")
n_q, n_items, n_users = int(sys.argv[1]), int(sys.argv[2]), 1_000_000
mask_per = int(sys.argv[3]) if len(sys.argv) > 3 else 100
g = torch.Generator(device=dev).manual_seed(1)
U = torch.randn(n_users, 64, device=dev, generator=g) * 0.125
I = torch.randn(n_items, 64, device=dev, generator=g) * 0.125
uids = torch.arange(n_q, device=dev, dtype=torch.int32)
mrp = mc = None
if mask_per:
    mrp = torch.arange(0, (n_q + 1) * mask_per, mask_per, device=dev, dtype=torch.int64)
    x = torch.sort(torch.randint(0, n_items - mask_per, (n_q, mask_per), device=dev, generator=g), dim=1).values
    mc = (x + torch.arange(mask_per, device=dev)).to(torch.int32).flatten().contiguous()
for _ in range(2):
    ops.score_topk(U, I, 20, user_ids=uids, mask_rowptr=mrp, mask_col=mc, precision=ops.SCORE_TF32_CHECKED)
torch.cuda.synchronize()
n_units = min(8192, (n_q + 255) // 256)
buf = (ctypes.c_ulonglong * (n_units * 16))()
_lib.check(_lib.load().cr_debug_tc_timeline(buf, n_units), "timeline")
t = np.frombuffer(buf, dtype=np.uint64).reshape(n_units, 16).astype(np.int64)
t0 = t[:, 0].min()
names = ["entry", "setup", "q->tmem", "tile0", "tile15", "tile127", "tile1023", "tile4095", "tile8191", "last", "exit"]
print(f"n_q={n_q} n_items={n_items} mask_per={mask_per} units={n_units} dbg={os.environ['CR_TC_DEBUG_MODE']}")
rel = (t[:, :11] - t[:, :1]) / 1e3          # us since the unit's own entry
for k, nm in enumerate(names):
    ok = t[:, k] > 0
    if ok.any():
        print(f"  {nm:9s} since unit entry: min {rel[ok, k].min():10.1f}  median {np.median(rel[ok, k]):10.1f}  max {rel[ok, k].max():10.1f} us")
print(f"  unit entry spread: {(t[:, 0].max() - t0) / 1e3:.1f} us; first entry -> last exit: {(t[:, 10].max() - t0) / 1e3:.1f} us")
seg = [(3, 4, 15), (4, 5, 112), (5, 6, 896), (6, 7, 3072), (7, 8, 4096)]
for a, b, n in seg:
    ok = (t[:, a] > 0) & (t[:, b] > 0)
    if ok.any():
        print(f"  us/tile between {names[a]} and {names[b]}: median {np.median((t[ok, b] - t[ok, a]) / 1e3 / n):.3f}")
span = np.median(t[:, 10] - t[:, 0]) / 1e3
for k, nm in ((11, "epilogue w0 blocked on tfull (MMA)"), (12, "epilogue w0 blocked on mfull (mask)"), (13, "MMA blocked on full (TMA)"),
              (14, "MMA blocked on tempty (epilogue)"), (15, "mask producer blocked on mempty (epilogue)")):
    cyc = np.median(t[:, k])
    print(f"  {nm:45s}: median {cyc / 1e3:10.0f} kcycles  (~{cyc / 1.9e3:8.0f} us of {span:.0f} us)")
