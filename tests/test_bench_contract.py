"""bench.py's one-line JSON contract: the CPU reference arm here (no GPU needed), the B200 arm at toy sizes on the GPU box."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e", "gpu_launches"}


def _run(args, env=None, timeout=600):
    e = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        e.pop(k, None)
    e.update(env or {})
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, env=e, timeout=timeout, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    return [json.loads(l) for l in p.stdout.splitlines() if l.startswith("{")]


def test_reference_arm_line_and_non_zero_ranks_stay_silent():
    small = ["--impl", "reference", "--steps", "2", "--warmup", "1", "--n-items", "30000", "--n-users", "4000", "--cpu-sample-users", "8",
             "--configs", "C2"]
    (line,) = _run(small)
    assert BASE_KEYS <= set(line) and line["impl"] == "reference"
    assert line["unit"] == "users/s" and line["higher_is_better"] is True and line["value"] > 0 and line["gpu_launches"] == 0
    assert line["steps"] == 2 and line["warmup"] == 1 and line["n_gpus"] == 1
    cb = line["cpu_baseline"]
    have_ref = os.path.isfile(os.path.join(ROOT, "baseline", "_ref", "model", "BaseRecommender.py"))
    # the reference's own _evaluate when baseline/_ref travelled with the snapshot, else the oracle port
    assert cb["kind"] == ("reference" if have_ref else "port") and cb["cores"] >= 1 and cb["value"] == line["value"]
    assert "8 users/step" in cb["sample"]
    c2 = line["extra"]["C2"]
    if have_ref:
        assert c2["cpu_baseline"]["kind"] == "reference" and c2["value"] > 0 and "LGCN_Encoder.forward" in c2["cpu_baseline"]["sample"]
    else:
        assert "unavailable" in c2["cpu_baseline"]
    assert line["e2e"] == {"value": line["value"], "unit": "users/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and "model" not in line["config"]
    # under torchrun only rank 0 runs the CPU arm; the other ranks exit 0 without a line
    assert _run(small, env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}) == []


@pytest.mark.gpu
@pytest.mark.timeout(600)
def test_b200_arm_line_at_toy_sizes():
    (line,) = _run(["--steps", "3", "--warmup", "3", "--n-items", "300000", "--n-users", "60000", "--users-per-step", "4096",
                    "--graph-edges", "3000000", "--cpu-sample-users", "16", "--train-batch", "1024", "--configs", "C1,C2", "--config-steps", "2"])
    assert BASE_KEYS <= set(line) and "impl" not in line
    assert line["unit"] == "users/s" and line["value"] > 0 and line["gpu_launches"] > 0 and line["vs_baseline"] is None
    assert line["e2e"]["value"] > 0 and line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0
    for part in (line, line["lightgcn"]):
        r = part["roofline"]
        assert r["bound"] in ("hbm", "tensor") and r["achieved"] > 0 and r["peak"] > 0 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3
        assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(part["clocks"])
        assert part["cpu_baseline"]["kind"] in ("port", "reference") and part["cpu_baseline"]["value"] > 0
        assert part["gpu_library_baseline"].get("value", 0) > 0, part["gpu_library_baseline"]
    assert line["lightgcn"]["unit"] == "edges/s" and line["lightgcn"]["gpu_launches"] > 0
    assert line["lightgcn"]["train_step"]["value"] > 0
    for key in ("C1", "C2"):
        x = line["extra"][key]
        assert "error" not in x, x
        assert x["value"] > 0 and x["gpu_launches"] > 0 and x["e2e"]["first_call_ms"] > 0 and x["roofline"]["kernel_launches_per_step"] > 0
        assert x["cpu_baseline"].get("value", 0) > 0 and x["gpu_library_baseline"].get("value", 0) > 0, x


def test_bench_config_datasets_follow_the_reference_split():
    """bench_configs: the synthetic C2-shaped data set and its array restatement of data/split.py + data/convert.py — every user and
    item keeps a record (the reference's id2item covers every row of the tables), warm records 8:1:1 with val / test users and items
    all present in train, cold items disjoint from train and split by item, overall = cold + warm over users present in both."""
    import numpy as np
    sys.path.insert(0, ROOT)
    import bench_configs as B
    c = B.CONFIGS["C2"]
    rng = np.random.default_rng(c["seed"])
    pairs = B.synth_pairs(rng, c["n_users"], c["n_items"], c["n_inter"], c["zipf"])
    assert len(pairs) == c["n_inter"] and len(np.unique(pairs[:, 0] * c["n_items"] + pairs[:, 1])) == c["n_inter"]
    assert len(np.unique(pairs[:, 0])) == c["n_users"] and len(np.unique(pairs[:, 1])) == c["n_items"]
    s, info = B.split_item_cold(rng, pairs, c["n_users"], c["n_items"])
    tr_u, tr_i = set(s["training"][:, 0].tolist()), set(s["training"][:, 1].tolist())
    for k in ("warm_valid", "warm_test"):
        assert set(s[k][:, 0].tolist()) <= tr_u and set(s[k][:, 1].tolist()) <= tr_i
    cold_items = set(info["cold_item"].tolist())
    assert not (cold_items & tr_i) and abs(len(cold_items) / c["n_items"] - 0.2) < 0.01
    assert not (set(s["cold_valid"][:, 1].tolist()) & set(s["cold_test"][:, 1].tolist()))
    for k, (cold, warm) in {"overall_valid": ("cold_valid", "warm_valid"), "overall_test": ("cold_test", "warm_test")}.items():
        both = set(s[cold][:, 0].tolist()) & set(s[warm][:, 0].tolist())
        assert set(s[k][:, 0].tolist()) == both
    n_warm = len(s["training"]) + len(s["warm_valid"]) + len(s["warm_test"])
    assert n_warm + len(s["cold_valid"]) + len(s["cold_test"]) == c["n_inter"]
    assert abs(len(s["training"]) / n_warm - 0.8) < 0.05
    data, _ = B.make_dataset("C2")
    assert data.training_size()[:2] == (c["n_users"], c["n_items"]) and len(data.id2item) == c["n_items"]
