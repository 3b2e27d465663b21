"""Probe: one d = 128 sweep at the headline shape (75,776 users x 10M items), TF32-checked vs the exact kernel on a user sample."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from coldrec_b200 import ops, _lib

dev = torch.device("cuda:0"); lib = _lib.load()
g = torch.Generator(device=dev).manual_seed(128)
n_users, n_items, n_q = 1_000_000, 10_000_000, 75_776
U = torch.randn(n_users, 128, device=dev, generator=g) * 0.09
I = torch.randn(n_items, 128, device=dev, generator=g) * 0.09
uids = torch.arange(n_q, device=dev, dtype=torch.int32)
for _ in range(2):
    s, i, nref = ops.score_topk(U, I, 20, user_ids=uids, precision=ops.SCORE_TF32_CHECKED)
torch.cuda.synchronize()
lib.cr_profile_enable(1)
s, i, nref = ops.score_topk(U, I, 20, user_ids=uids, precision=ops.SCORE_TF32_CHECKED)
torch.cuda.synchronize()
tot, cnt = ctypes.c_double(), ctypes.c_int(); lib.cr_profile_read(0, ctypes.byref(tot), ctypes.byref(cnt))
se, ie, _ = ops.score_topk(U, I, 20, user_ids=uids[:512].contiguous(), precision=ops.SCORE_EXACT_F32)
print(json.dumps({"sweep_ms": round(tot.value / max(cnt.value, 1), 3), "tflops": round(2.0 * n_q * n_items * 128 / (tot.value / max(cnt.value, 1)) / 1e9, 1),
                  "n_refined": int(nref.item()), "ids_equal_exact_on_512_users": bool(torch.equal(ie, i[:512])), "scores_equal": bool(torch.equal(se, s[:512]))}))
