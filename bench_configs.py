"""BASELINE.json configs[0..2] (C1 BPR-MF / MovieLens-shaped, C2 LightGCN / CiteULike-shaped, C3 DropoutNet + Heater /
XING-shaped) as bench lines: what a ColdRec user runs at the end of ``run()`` — test + evaluate under all / cold / warm —
through the trainer API with HOST dicts in and metric strings out.

Per config, three arms on the same seeded synthetic data (``ArrayDataBuilder`` serves all three: it exposes the
reference's dict views next to its arrays):
  fused                 ``coldrec_b200.BaseColdStartTrainer``: ``tr.test(kind)`` + ``full_evaluation`` (+ K3 propagation for
                        C2, + K4 towers for C3).  ``first_call`` includes building the eval plans from the host data
                        (``_get_eval_cache``: memoised afterwards, exactly as in the reference); ``value`` is the steady state.
  gpu_library_baseline  the reference's own classes from baseline/_ref on the same GPU (its torch path, ``--use_gpu true``).
  cpu_baseline          the same reference classes on the host cores (``kind: "reference"``; the oracle port if
                        baseline/_ref is absent, ``kind: "port"``).
These workloads are launch / latency bound on a B200 (a 6,040 x 3,706 score matrix is 2.9 GFLOP): the honest roofline
figure is the number of kernel launches per step and the device time against the step's wall time; both are reported.
"""
import contextlib
import io
import os
import sys
import time
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
D, TOPN = 64, "10,20"

CONFIGS = {
    "C1": dict(name="BPR-MF item cold-start, MovieLens-shaped", n_users=6040, n_items=3706, n_inter=1_000_209, content=206, zipf=1.0, seed=2024),
    "C2": dict(name="LightGCN 3-layer propagation + full ranking, CiteULike-shaped", n_users=5551, n_items=16980, n_inter=204_986,
               content=300, zipf=0.8, seed=3, layers=3),
    "C3": dict(name="DropoutNet / Heater cold-item generation + cold-item top-20, XING-shaped", n_users=106_881, n_items=20_519,
               n_inter=3_856_580, content=2738, zipf=0.8, seed=4),
}


# ------------------------------------------------------------------------------------------- synthetic data
def synth_pairs(rng, n_users, n_items, n_inter, zipf):
    """Unique (user, item) pairs: lognormal user activity, Zipf item popularity, every user and item present."""
    pu = rng.lognormal(4.6, 1.0, n_users); pu /= pu.sum()
    pi = 1.0 / np.arange(1, n_items + 1) ** zipf; pi = rng.permutation(pi); pi /= pi.sum()
    base = np.unique(np.concatenate([np.arange(n_users, dtype=np.int64) * n_items + rng.integers(0, n_items, n_users),
                                     rng.integers(0, n_users, n_items).astype(np.int64) * n_items + np.arange(n_items)]))
    keys = base
    while len(keys) < n_inter:           # the popular cells collide a lot: over-draw generously, de-duplicate once per round
        m = max(3 * (n_inter - len(keys)), 200_000)
        keys = np.union1d(keys, rng.choice(n_users, m, p=pu).astype(np.int64) * n_items + rng.choice(n_items, m, p=pi))
    extra = np.setdiff1d(keys, base)     # the guaranteed pairs always stay: every user and every item keeps a record
    keys = np.concatenate([base, extra[rng.permutation(len(extra))[:n_inter - len(base)]]])
    keys = keys[rng.permutation(len(keys))]
    return np.stack([keys // n_items, keys % n_items], 1)


def split_item_cold(rng, pairs, n_users, n_items):
    """Array restatement of data/split.py + data/convert.py (--cold_object item, defaults): 80 % of the items warm, their
    records 8:1:1 (val / test records whose user or item is missing from train move to train), the cold items split 50/50
    into val / test by item, overall = cold + warm restricted to users present in both."""
    items = rng.permutation(np.unique(pairs[:, 1]))
    warm_item = np.zeros(n_items, bool); warm_item[items[:int(0.8 * len(items))]] = True
    warm = pairs[warm_item[pairs[:, 1]]]
    cold = pairs[~warm_item[pairs[:, 1]]]
    warm = warm[rng.permutation(len(warm))]
    n_val = n_test = int(0.1 * len(warm))
    train, val, test = warm[:len(warm) - n_val - n_test], warm[len(warm) - n_val - n_test:len(warm) - n_test], warm[len(warm) - n_test:]
    for _ in range(2):
        for col, n in ((0, n_users), (1, n_items)):
            seen = np.zeros(n, bool); seen[train[:, col]] = True
            mv = ~seen[val[:, col]]
            train, val = np.concatenate([train, val[mv]]), val[~mv]
            seen[train[:, col]] = True
            mt = ~seen[test[:, col]]
            train, test = np.concatenate([train, test[mt]]), test[~mt]
    cold_items = rng.permutation(np.unique(cold[:, 1]))
    in_val = np.zeros(n_items, bool); in_val[cold_items[:len(cold_items) // 2]] = True
    cold_val, cold_test = cold[in_val[cold[:, 1]]], cold[~in_val[cold[:, 1]]]

    def overall(c, w):
        both = np.intersect1d(c[:, 0], w[:, 0])
        o = np.concatenate([c, w])
        return o[np.isin(o[:, 0], both)]
    splits = dict(training=train, warm_valid=val, cold_valid=cold_val, overall_valid=overall(cold_val, val), warm_test=test,
                  cold_test=cold_test, overall_test=overall(cold_test, test))
    info = dict(warm_user=np.unique(train[:, 0]), warm_item=np.unique(train[:, 1]), cold_user=np.unique(cold[:, 0]),
                cold_item=np.unique(cold[:, 1]))
    return splits, info


def make_dataset(key):
    from coldrec_b200 import ArrayDataBuilder
    c = CONFIGS[key]
    rng = np.random.default_rng(c["seed"])
    pairs = synth_pairs(rng, c["n_users"], c["n_items"], c["n_inter"], c["zipf"])
    s, info = split_item_cold(rng, pairs, c["n_users"], c["n_items"])
    if key == "C3":      # sparse-ish like XING
        content = ((rng.random((c["n_items"], c["content"]), dtype=np.float32) < 0.02) *
                   rng.standard_normal((c["n_items"], c["content"]), dtype=np.float32))
    else:
        content = rng.standard_normal((c["n_items"], c["content"]), dtype=np.float32)
    data = ArrayDataBuilder(s["training"], s["warm_valid"], s["cold_valid"], s["overall_valid"], s["warm_test"], s["cold_test"],
                            s["overall_test"], c["n_users"], c["n_items"], info["warm_user"], info["warm_item"], info["cold_user"],
                            info["cold_item"], None, content)
    return data, c


def _args(model):
    return types.SimpleNamespace(topN=TOPN, model=model, dataset="syn", emb_size=D, epochs=0, bs=4096, lr=1e-3, reg=1e-4, early_stop=0,
                                 eval_every=1, cold_object="item", save_emb=False, layers=3, seed=1)


class _Cfg:
    def __init__(self, args, data, device):
        self.args, self.data, self.device = args, data, torch.device(device)


@contextlib.contextmanager
def _quiet():
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        yield buf


def _tables(c, seed, device="cpu"):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(c["n_users"], D, generator=g) * 0.1).to(device), (torch.randn(c["n_items"], D, generator=g) * 0.1).to(device)


def _tower_states(c, seed):
    """Seeded weights in the reference modules' own state_dict layout (DeepCF: [V | content] -> 200 -> 100 -> 64 with eval
    BatchNorm; Heater: gate + shared expert MLP + blend + out_linear + final_trans)."""
    g = torch.Generator().manual_seed(seed)
    C = c["content"]
    lin = lambda o, k, std: (torch.randn(o, k, generator=g) * std, torch.randn(o, generator=g) * 0.05)
    dn = {}
    for side, k0 in (("u", D), ("v", D + C)):
        dims = [k0, 200, 100]
        for l in range(2):
            w, b = lin(dims[l + 1], dims[l], 0.08 if dims[l] < 1000 else 0.03)
            dn.update({f"{side}_layers.{l}.layer.weight": w, f"{side}_layers.{l}.layer.bias": b,
                       f"{side}_layers.{l}.bn.weight": torch.rand(dims[l + 1], generator=g) + 0.5,
                       f"{side}_layers.{l}.bn.bias": torch.randn(dims[l + 1], generator=g) * 0.1,
                       f"{side}_layers.{l}.bn.running_mean": torch.randn(dims[l + 1], generator=g) * 0.05,
                       f"{side}_layers.{l}.bn.running_var": torch.rand(dims[l + 1], generator=g) * 0.5 + 0.5,
                       f"{side}_layers.{l}.bn.num_batches_tracked": torch.tensor(1)})
        w, b = lin(D, 100, 0.1)
        dn.update({f"{side}_emb.weight": w, f"{side}_emb.bias": b})
    ht = {}
    for name, (o, k, std) in {"gate.linear": (5, C, 0.03), "fc.linear1": (200, C, 0.03), "fc.linear2": (D, 200, 0.08),
                              "out_linear": (D, D, 0.1), "final_trans": (D, D, 0.1)}.items():
        ht[name + ".weight"], ht[name + ".bias"] = lin(o, k, std)
    return dn, ht


# ------------------------------------------------------------------------------------------- arms
def _eval_three(tr):
    """The tail of run() (model/BaseRecommender.py:363-370): test + evaluate under all / cold / warm.  Returns eval-user count."""
    n = 0
    with _quiet():
        for kind in ("all", "cold", "warm"):
            rec = tr.test(test_type=kind)
            tr.full_evaluation(rec, test_type=kind)
            n += len(rec)
    return n


def _timed(fn, steps, warmup, cuda):
    for _ in range(warmup):
        fn()
    if cuda:
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        out = fn()
    if cuda:
        torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps, out


def _fused_trainer(data, device, tables=None):
    import coldrec_b200 as cr

    class T(cr.BaseColdStartTrainer):
        def train(self): pass
        def save(self): pass
    tr = T(_Cfg(_args("MF"), data, device))
    if tables is not None:
        tr.user_emb, tr.item_emb = tables
    return tr


def _reference_modules():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    try:
        import refimport
    finally:
        sys.path.pop(0)
    if not refimport.available():
        return None
    refimport.import_reference()
    from model.MF import MF
    from model.LightGCN import LGCN_Encoder
    from model.DropoutNet import get_model
    from model.Heater import Heater_encoder
    return types.SimpleNamespace(MF=MF, LGCN_Encoder=LGCN_Encoder, get_model=get_model, Heater_encoder=Heater_encoder)


def _reference_trainer(ref, data, device, tables):
    """The reference's own MF trainer class (``_evaluate`` / ``batch_predict`` / ``full_evaluation`` untouched) holding given tables."""
    with _quiet():
        tr = ref.MF(_Cfg(_args("MF"), data, device))
    tr.user_emb, tr.item_emb = tables[0].to(device), tables[1].to(device)
    return tr


def _ref_towers(ref, c, dn, ht, U, V, content, device):
    enc = ref.get_model(D, 0, c["content"], [200, 100], D).to(device).eval()
    enc.load_state_dict({k: v.to(device) for k, v in dn.items()})
    hen = ref.Heater_encoder(D, 0, c["content"], [200, D], D, 1e-4, 5, 0.5).to(device).eval()
    hen.load_state_dict({k: v.to(device) for k, v in ht.items()})
    U, V, content = U.to(device), V.to(device), content.to(device)

    def run():
        with torch.no_grad():
            du, dv = enc.encode(U, V, None, content)
            hu, hv, _, _ = hen.encode(U, V, None, content)
        return du, dv, hu, hv
    return run


def run_config(key, device, lib, steps=5, warmup=2, cpu=True, host_threads=None):
    """One bench line for C1 / C2 / C3 (N = 1)."""
    import coldrec_b200 as cr
    from coldrec_b200 import towers
    data, c = make_dataset(key)
    line = {"workload": f"{key} {c['name']}: {c['n_users']} users x {c['n_items']} items, {c['n_inter']} interactions, d={D}, "
                        f"item cold-start, test + evaluate under all / cold / warm (the tail of run())", "data": "synthetic"}
    U, V = _tables(c, c["seed"] + 1)
    content = torch.from_numpy(np.asarray(data.mapped_item_content, dtype=np.float32))
    ref = _reference_modules()
    extra_fused = {}

    # ---- fused arm ------------------------------------------------------------------------------------------------
    t0 = time.perf_counter()
    tr = _fused_trainer(data, device)
    if key == "C1":
        tr.user_emb, tr.item_emb = U.to(device), V.to(device)
        step = lambda: _eval_three(tr)
    elif key == "C2":
        G = data.graph(device)
        Ud, Vd = U.to(device), V.to(device)

        def step():
            tr.user_emb, tr.item_emb = cr.propagate(G, Ud, Vd, c["layers"])       # LGCN_Encoder.forward at eval time
            return _eval_three(tr)
    else:
        dn, ht = _tower_states(c, 11)
        dnd = {k: v.to(device) for k, v in dn.items() if v.dtype == torch.float32}
        htd = {k: v.to(device) for k, v in ht.items()}
        Ud, Vd = U.to(device), V.to(device)
        Cd = towers.prepare(content.to(device))       # the constant content table in the layer kernel's operand form, once (as a trainer would)

        def step():
            n = 0
            for gen in (lambda: towers.dropoutnet_encode(dnd, Ud, Vd, None, Cd), lambda: towers.heater_encode(htd, Ud, Vd, Cd, 5, 0.5)):
                tr.user_emb, tr.item_emb = gen()
                with _quiet():
                    rec = tr.test(test_type="cold")
                    tr.full_evaluation(rec, test_type="cold")
                n += len(rec)
            return n
        # K4 alone: both towers over all 20,519 items (content read 225 MB) and all users
        k4_s, _ = _timed(lambda: (towers.dropoutnet_encode(dnd, Ud, Vd, None, Cd), towers.heater_encode(htd, Ud, Vd, Cd, 5, 0.5)), steps, warmup, True)
        flop = 2.0 * c["n_items"] * ((D + c["content"]) * 200 + 200 * 100 + 100 * D + c["content"] * (200 + 5) + 200 * D + 2 * D * D) \
            + 2.0 * c["n_users"] * (D * 200 + 200 * 100 + 100 * D + 2 * D * D)
        extra_fused["k4_towers"] = {"ms": round(k4_s * 1e3, 3), "tflops": round(flop / k4_s / 1e12, 2), "flop": flop,
                                    "content_bytes": int(content.numel() * 4 * 2),
                                    "content_gbs": round(content.numel() * 4 * 2 / k4_s / 1e9, 1),
                                    "what": "DropoutNet (item + user towers) and Heater encode over every item / user, eval mode"}
    n_eval = step()
    torch.cuda.synchronize()
    first_s = time.perf_counter() - t0
    l0 = lib.cr_launch_count()
    dt, _ = _timed(step, steps, warmup, True)
    launches = (lib.cr_launch_count() - l0) // (steps + warmup)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    line.update(metric="eval users/sec through the trainer API (host dicts in, metric strings out)", unit="users/s",
                value=round(n_eval / dt, 1), ms_per_step=round(dt * 1e3, 3), eval_users_per_step=n_eval, steps=steps, warmup=warmup,
                e2e={"value": round(n_eval / dt, 1), "unit": "users/s", "first_call_ms": round(first_s * 1e3, 1),
                     "first_call_includes": "trainer construction, item flags, EvalPlan build from the host data (memoised afterwards, "
                                            "as the reference memoises _get_eval_cache), first launches",
                     "d2h": "metric partial sums per setting (the per-step sync); rec lists stay lazy on the device"},
                gpu_launches=int(launches), **extra_fused)
    # device time of one step vs its wall time: how launch / host bound the step is
    torch.cuda.synchronize()
    ev0.record(); step(); ev1.record(); torch.cuda.synchronize()
    if key == "C1":
        work = {"flop": 2.0 * n_eval * c["n_items"] * D}
    elif key == "C2":
        nnz = data.graph(device).nnz
        work = {"flop": 2.0 * n_eval * c["n_items"] * D, "spmm_bytes": c["layers"] * (nnz * (8 + 4 * D) + 3 * (c["n_users"] + c["n_items"]) * 4 * D)}
    else:
        work = {"flop": 2.0 * n_eval * c["n_items"] * D + 2 * extra_fused["k4_towers"]["flop"]}
    line["roofline"] = {"bound": "latency (launch / host bound: the whole step is %.1f GFLOP)" % (work["flop"] / 1e9),
                        "kernel_launches_per_step": int(launches), "step_ms_device_events": round(ev0.elapsed_time(ev1), 3),
                        "step_ms_wall": round(dt * 1e3, 3), "achieved": round(work["flop"] / dt / 1e12, 3), "unit": "TFLOP/s",
                        "peak": None, "frac": None, **{k: v for k, v in work.items() if k != "flop"}}

    # ---- the reference's own classes: same GPU (torch path), then host cores ---------------------------------------
    tw = _tower_states(c, 11) if key == "C3" else (None, None)
    ref_step_for = lambda dev: reference_step(ref, key, data, c, U, V, content, tw[0], tw[1], dev)

    if ref is not None:
        try:
            s = ref_step_for(str(device))
            ldt, ln = _timed(s, 1, 1, True)
            line["gpu_library_baseline"] = {"value": round(ln / ldt, 1), "unit": "users/s", "ms_per_step": round(ldt * 1e3, 1),
                                            "kind": "reference classes from baseline/_ref on this GPU (torch matmul + per-user index_put + "
                                                    "topk + host ranking_evaluation" + ("; torch.sparse.mm COO" if key == "C2" else "") +
                                                    ("; nn.Linear/BatchNorm1d towers" if key == "C3" else "") + ")"}
        except Exception as ex:
            line["gpu_library_baseline"] = {"error": f"{type(ex).__name__}: {str(ex)[:160]}"}
        torch.cuda.empty_cache()
    if cpu:
        line["cpu_baseline"] = cpu_config_baseline(key, data, c, U, V, content, ref_step_for, host_threads)
    return line


def reference_step(ref, key, data, c, U, V, content, dn, ht, dev):
    """One step of a config on the reference's own classes (baseline/_ref) on device ``dev``; None without baseline/_ref."""
    if ref is None:
        return None
    rt = _reference_trainer(ref, data, dev, (U, V))
    if key == "C1":
        return lambda: _eval_three(rt)
    if key == "C2":
        with _quiet():
            enc = ref.LGCN_Encoder(data, D, c["layers"], torch.device(dev))
        enc.embedding_dict["user_emb"].data, enc.embedding_dict["item_emb"].data = U.to(dev), V.to(dev)
        enc = enc.to(dev)

        def s():
            with torch.no_grad():
                rt.user_emb, rt.item_emb = enc()
            return _eval_three(rt)
        return s
    gen = _ref_towers(ref, c, dn, ht, U, V, content, dev)

    def s():
        n = 0
        du, dv, hu, hv = gen()
        for ue, ie in ((du, dv), (hu, hv)):
            rt.user_emb, rt.item_emb = ue, ie
            with _quiet():
                rec = rt.test(test_type="cold")
                rt.full_evaluation(rec, test_type="cold")
            n += len(rec)
        return n
    return s


def reference_config_line(key, cores):
    """`bench.py --impl reference`: the C1 / C2 / C3 step on the reference's own classes, host cores."""
    data, c = make_dataset(key)
    U, V = _tables(c, c["seed"] + 1)
    content = torch.from_numpy(np.asarray(data.mapped_item_content, dtype=np.float32))
    tw = _tower_states(c, 11) if key == "C3" else (None, None)
    ref = _reference_modules()
    base = cpu_config_baseline(key, data, c, U, V, content, lambda dev: reference_step(ref, key, data, c, U, V, content, tw[0], tw[1], dev), cores)
    return {"workload": f"{key} {c['name']}", "impl": "reference", "metric": "eval users/sec through the trainer API (host dicts in, metric strings out)",
            "value": base.get("value"), "unit": "users/s", "cpu_baseline": base}


def cpu_config_baseline(key, data, c, U, V, content, ref_step_for, host_threads=None):
    """The reference's own code on the host cores (kind "reference"); without baseline/_ref the oracle port of _evaluate."""
    torch.set_num_threads(host_threads or (len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()))
    s = ref_step_for("cpu")
    if s is None:
        return {"unavailable": "baseline/_ref missing (python baseline/install_ref.py)", "kind": "port"}
    t0 = time.perf_counter()
    n = s()
    dt = time.perf_counter() - t0
    return {"value": round(n / dt, 1), "unit": "users/s", "cores": torch.get_num_threads(), "kind": "reference",
            "sample": f"full size, one step: {n} eval users, the reference's own MF._evaluate + util.evaluator.ranking_evaluation"
                      + (" + LGCN_Encoder.forward" if key == "C2" else "") + (" + DeepCF / Heater_encoder encode" if key == "C3" else "")
                      + f" (baseline/_ref), {dt:.1f} s"}
