"""Shared fixtures: load tests/golden/*.npz (outputs of the real reference, see oracle/make_golden.py)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))


def splits_of(g):
    rows = lambda a: [[int(u), int(i), 1.0] for u, i in a]
    return {k[len("split_"):]: rows(v) for k, v in g.items() if k.startswith("split_")}


def builder_args(g, with_content=True):
    """Positional args of ColdStartDataBuilder (util/databuilder.py:7-10) from a golden file."""
    s = splits_of(g)
    cold_object = str(g["cold_object"])
    uc = g["content"] if (with_content and cold_object == "user") else None
    ic = g["content"] if (with_content and cold_object == "item") else None
    return (s["training"], s["warm_valid"], s["cold_valid"], s["overall_valid"], s["warm_test"], s["cold_test"],
            s["overall_test"], int(g["info_user_num"]), int(g["info_item_num"]), g["info_warm_user"].tolist(),
            g["info_warm_item"].tolist(), g["info_cold_user"].tolist(), g["info_cold_item"].tolist(), uc, ic)


def rec_from_golden(g, prefix):
    users = g[f"{prefix}_users"].tolist()
    return {u: list(zip(g[f"{prefix}_raw_ids"][j].tolist(), g[f"{prefix}_scores"][j])) for j, u in enumerate(users)}


def t(x):
    return torch.from_numpy(np.ascontiguousarray(x))
