#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (ColdRec at /root/reference).

TEST INFRASTRUCTURE ONLY.  Run in the build container (the reference tree is not on the GPU box):

    python oracle/make_golden.py            # writes tests/golden/{eval_item,eval_user,eval_tiny,graph,towers,train}.npz
    python oracle/make_golden.py --only train   # just one of them

What is executed from the reference, unmodified: ``data/split.py`` + ``data/convert.py`` (subprocess),
``util.loader.DataLoader``, ``util.databuilder.ColdStartDataBuilder`` / ``TorchGraphInterface``,
``model.BaseRecommender.BaseColdStartTrainer._evaluate`` (through ``test()``), the ``batch_predict`` of
``model.MF`` / ``model.ALDI`` / ``model.VBPR``, ``util.evaluator.ranking_evaluation``,
``model.LightGCN.LGCN_Encoder``, ``model.SimGCL.SimGCL_Encoder``, ``model.NGCF.NGCF_Encoder``,
``model.DropoutNet.DropoutNet_Learner``, ``model.Heater.Heater_Learner``, ``model.GAR.GAR_Learner``,
``model.ALDI.ALDI_Learner``; for train.npz the training loop body of ``model/LightGCN.py:21-28`` / ``model/MF.py:19-27``
(``util.utils.next_batch_pairwise``, ``bpr_loss``, ``l2_reg_loss``, ``torch.optim.Adam``).  A stub ``faiss`` module is injected because ``model/__init__.py``
imports KNN/NCL which import faiss (absent here).  Inputs are seeded; fixtures stay small.
"""
import argparse
import os
import pickle
import subprocess
import sys
import tempfile
import types

import numpy as np
import torch

REF = os.environ.get("COLDREC_REF", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
K, TOPN, D = 20, "10,20", 64


def _import_reference():
    sys.modules.setdefault("faiss", types.ModuleType("faiss"))
    sys.path.insert(0, REF)


def synth_interactions(rng, n_users, n_items, n_inter):
    """Unique (user, item) pairs; lognormal user activity, Zipf-ish item popularity; every id used."""
    pu = rng.lognormal(0.0, 1.0, n_users); pu /= pu.sum()
    pi = 1.0 / np.arange(1, n_items + 1) ** 0.8; pi = rng.permutation(pi); pi /= pi.sum()
    pairs = set()
    for u in range(n_users):                       # everyone appears at least a few times
        for it in rng.choice(n_items, size=6, replace=False, p=pi):
            pairs.add((u, int(it)))
    for it in range(n_items):
        for u in rng.choice(n_users, size=3, replace=False, p=pu):
            pairs.add((int(u), it))
    while len(pairs) < n_inter:
        us = rng.choice(n_users, size=n_inter, p=pu)
        its = rng.choice(n_items, size=n_inter, p=pi)
        for u, it in zip(us, its):
            pairs.add((int(u), int(it)))
            if len(pairs) >= n_inter:
                break
    arr = np.array(sorted(pairs), dtype=np.int64)
    return arr[rng.permutation(len(arr))]


def prepare_dataset(work, name, cold_object, rng, n_users, n_items, n_inter, content_dim):
    """Write the reference's on-disk format and run its own split/convert scripts."""
    ddir = os.path.join(work, "data", name)
    os.makedirs(ddir, exist_ok=True)
    inter = synth_interactions(rng, n_users, n_items, n_inter)
    with open(os.path.join(ddir, f"{name}.csv"), "w") as f:
        f.write("user,item\n")
        for u, i in inter:
            f.write(f"{u},{i}\n")
    n_obj = n_items if cold_object == "item" else n_users
    content = rng.standard_normal((n_obj, content_dim)).astype(np.float32)
    np.save(os.path.join(ddir, f"{name}_{cold_object}_content.npy"), content)
    for script in ("split.py", "convert.py"):
        subprocess.run([sys.executable, os.path.join(REF, "data", script), "--dataset", name,
                        "--cold_object", cold_object, "--datadir", "./"],
                       cwd=os.path.join(work, "data"), check=True, stdout=subprocess.DEVNULL)
    return content


def load_config(work, name, cold_object):
    """The body of main.py:23-58 (Config) without argparse: reference loaders + builder."""
    from util.loader import DataLoader
    from util.databuilder import ColdStartDataBuilder
    base = os.path.join(work, "data", name, f"cold_{cold_object}")
    L = lambda f: DataLoader.load_data_set(os.path.join(base, f))
    splits = dict(training=L("warm_train.csv"), overall_valid=L("overall_val.csv"), warm_valid=L("warm_val.csv"),
                  cold_valid=L(f"cold_{cold_object}_val.csv"), overall_test=L("overall_test.csv"),
                  warm_test=L("warm_test.csv"), cold_test=L(f"cold_{cold_object}_test.csv"))
    info = pickle.load(open(os.path.join(base, "info_dict.pkl"), "rb"))
    content = np.load(os.path.join(work, "data", name, f"{name}_{cold_object}_content.npy"))
    uc, ic = (content, None) if cold_object == "user" else (None, content)
    data = ColdStartDataBuilder(splits["training"], splits["warm_valid"], splits["cold_valid"], splits["overall_valid"],
                                splits["warm_test"], splits["cold_test"], splits["overall_test"],
                                info["user_num"], info["item_num"], info["warm_user"], info["warm_item"],
                                info["cold_user"], info["cold_item"], uc, ic)
    return data, splits, info, content


def make_args(name, cold_object, model="MF", bs=32, **extra):
    ns = argparse.Namespace(topN=TOPN, model=model, dataset=name, emb_size=D, epochs=0, bs=bs, lr=1e-3, reg=1e-4,
                            early_stop=0, eval_every=1, cold_object=cold_object, save_emb=False, backbone="MF",
                            layers=3, **extra)
    return ns


class _Cfg:
    def __init__(self, args, data):
        self.args, self.data, self.device = args, data, torch.device("cpu")


def _pack_splits(splits):
    return {f"split_{k}": np.array([[r[0], r[1]] for r in v], dtype=np.int64).reshape(-1, 2) for k, v in splits.items()}


def _pack_info(info):
    return {f"info_{k}": np.asarray(info[k]) for k in ("user_num", "item_num", "warm_user", "warm_item", "cold_user", "cold_item")}


def _pack_rec(prefix, gt_set, rec_list, measure, perf, data):
    users = list(gt_set.keys())
    raw_ids = np.array([[it for it, _ in rec_list[u]] for u in users], dtype=np.int64)
    scores = np.array([[s for _, s in rec_list[u]] for u in users], dtype=np.float32)
    return {f"{prefix}_users": np.array(users, dtype=np.int64), f"{prefix}_raw_ids": raw_ids,
            f"{prefix}_dense_ids": np.vectorize(data.item.get)(raw_ids).astype(np.int64),
            f"{prefix}_scores": scores, f"{prefix}_measure": np.array(measure),
            f"{prefix}_performance": np.array(perf, dtype=np.float64)}


def golden_eval(work, name, cold_object, seed, n_users, n_items, n_inter, content_dim, variants=True):
    from model.BaseRecommender import BaseColdStartTrainer
    from model.MF import MF
    from model.ALDI import ALDI
    from model.VBPR import VBPR
    from util.evaluator import ranking_evaluation
    rng = np.random.default_rng(seed)
    prepare_dataset(work, name, cold_object, rng, n_users, n_items, n_inter, content_dim)
    data, splits, info, content = load_config(work, name, cold_object)
    g = torch.Generator().manual_seed(seed)
    user_emb = torch.randn(data.user_num, D, generator=g) * 0.3
    item_emb = torch.randn(data.item_num, D, generator=g) * 0.3

    class Stub(BaseColdStartTrainer):           # real _evaluate/test; batch_predict borrowed from MF
        train = save = predict = lambda self, *a: None
        batch_predict = MF.batch_predict
    t = Stub(_Cfg(make_args(name, cold_object), data))
    t.user_emb, t.item_emb = user_emb, item_emb
    out = dict(user_emb=user_emb.numpy(), item_emb=item_emb.numpy(), content=content, cold_object=np.array(cold_object),
               batch_size=np.array(t.batch_size), id2user=np.array([data.id2user[i] for i in range(len(data.user))]),
               id2item=np.array([data.id2item[i] for i in range(len(data.item))]),
               mapped_cold_item_idx=np.asarray(data.mapped_cold_item_idx), mapped_warm_item_idx=np.asarray(data.mapped_warm_item_idx),
               mapped_cold_user_idx=np.asarray(data.mapped_cold_user_idx), mapped_warm_user_idx=np.asarray(data.mapped_warm_user_idx))
    out.update(_pack_splits(splits)); out.update(_pack_info(info))
    for typ in ("all", "cold", "warm"):
        rec = t.test(typ)
        gt = {"all": data.overall_test_set, "cold": data.cold_test_set, "warm": data.warm_test_set}[typ]
        measure, perf = ranking_evaluation(gt, rec, t.topN)
        out.update(_pack_rec(f"mf_test_{typ}", gt, rec, measure, perf, data))
    rec = t.valid("all")
    measure, perf = ranking_evaluation(data.overall_valid_set, rec, [t.max_N])     # fast_evaluation's call, :293
    out.update(_pack_rec("mf_valid_all", data.overall_valid_set, rec, measure, perf, data))

    if variants and cold_object == "item":
        class StubALDI(Stub):
            batch_predict = ALDI.batch_predict
        a = StubALDI(_Cfg(make_args(name, cold_object, model="ALDI"), data))
        a.warm_user_emb, a.item_emb = user_emb, item_emb
        a.cold_user_emb = torch.randn(data.user_num, D, generator=g) * 0.3
        out["aldi_cold_user_emb"] = a.cold_user_emb.numpy()
        for typ in ("all", "cold"):
            rec = a.test(typ)
            gt = {"all": data.overall_test_set, "cold": data.cold_test_set}[typ]
            measure, perf = ranking_evaluation(gt, rec, a.topN)
            out.update(_pack_rec(f"aldi_test_{typ}", gt, rec, measure, perf, data))

        class StubVBPR(Stub):
            batch_predict = VBPR.batch_predict
        v = StubVBPR(_Cfg(make_args(name, cold_object, model="VBPR"), data))
        v.user_emb_main, v.item_emb_main = user_emb, item_emb
        v.user_emb_aux = torch.randn(data.user_num, D, generator=g) * 0.2
        v.item_emb_aux = torch.randn(data.item_num, D, generator=g) * 0.2
        out["vbpr_user_aux"], out["vbpr_item_aux"] = v.user_emb_aux.numpy(), v.item_emb_aux.numpy()
        rec = v.test("all")
        measure, perf = ranking_evaluation(data.overall_test_set, rec, v.topN)
        out.update(_pack_rec("vbpr_test_all", data.overall_test_set, rec, measure, perf, data))
    return out, data


def golden_graph(data, seed):
    from model.LightGCN import LGCN_Encoder
    from model.SimGCL import SimGCL_Encoder
    from model.NGCF import NGCF_Encoder
    adj = data.norm_adj.tocsr()
    adj.sort_indices()
    ui = data.ui_adj.tocsr(); ui.sort_indices()
    out = dict(user_num=np.array(data.user_num), item_num=np.array(data.item_num),
               train_u=np.array([data.user[p[0]] for p in data.training_data], dtype=np.int64),
               train_i=np.array([data.item[p[1]] for p in data.training_data], dtype=np.int64),
               adj_indptr=adj.indptr.astype(np.int64), adj_indices=adj.indices.astype(np.int64), adj_data=adj.data.astype(np.float32),
               ui_indptr=ui.indptr.astype(np.int64), ui_indices=ui.indices.astype(np.int64), ui_data=ui.data.astype(np.float32))
    torch.manual_seed(seed)
    with torch.no_grad():
        for L in (1, 2, 3):
            enc = LGCN_Encoder(data, D, L, torch.device("cpu"))          # xavier_uniform init, LightGCN.py:79-83
            if L == 1:
                out["E0_user"] = enc.embedding_dict["user_emb"].detach().numpy().copy()
                out["E0_item"] = enc.embedding_dict["item_emb"].detach().numpy().copy()
            else:
                enc.embedding_dict["user_emb"].copy_(torch.from_numpy(out["E0_user"]))
                enc.embedding_dict["item_emb"].copy_(torch.from_numpy(out["E0_item"]))
            u, i = enc.forward()
            out[f"lgcn_L{L}_user"], out[f"lgcn_L{L}_item"] = u.numpy().copy(), i.numpy().copy()
        s = SimGCL_Encoder(argparse.Namespace(eps=0.1), data, D, 3, torch.device("cpu"))
        s.embedding_dict["user_emb"].copy_(torch.from_numpy(out["E0_user"]))
        s.embedding_dict["item_emb"].copy_(torch.from_numpy(out["E0_item"]))
        u, i = s.forward(perturbed=False)
        out["simgcl_L3_user"], out["simgcl_L3_item"] = u.numpy().copy(), i.numpy().copy()
        n = NGCF_Encoder(data, D, 2, torch.device("cpu"))
        n.embedding_dict["user_emb"].copy_(torch.from_numpy(out["E0_user"]))
        n.embedding_dict["item_emb"].copy_(torch.from_numpy(out["E0_item"]))
        for l in range(2):
            out[f"ngcf_Wgc{l}_w"], out[f"ngcf_Wgc{l}_b"] = n.W_gc[l].weight.numpy().copy(), n.W_gc[l].bias.numpy().copy()
            out[f"ngcf_Wbi{l}_w"], out[f"ngcf_Wbi{l}_b"] = n.W_bi[l].weight.numpy().copy(), n.W_bi[l].bias.numpy().copy()
        u, i = n.forward()
        out["ngcf_L2_user"], out["ngcf_L2_item"] = u.numpy().copy(), i.numpy().copy()
    return out


def golden_train(data, seed, n_steps=3, bs=256, lr=5e-3, reg=1e-2):
    """The loop body of LightGCN.train / MF.train (model/LightGCN.py:21-28, model/MF.py:19-27) for a few batches:
    the reference's own sampler, encoder, losses and torch.optim.Adam.  lr / reg are larger than the CLI defaults
    so that three steps move the tables well above fp32 noise."""
    from model.LightGCN import LGCN_Encoder
    from model.MF import Matrix_Factorization
    from util.utils import next_batch_pairwise, bpr_loss, l2_reg_loss
    out = dict(lr=np.array(lr), reg=np.array(reg), bs=np.array(bs), n_steps=np.array(n_steps),
               train_u=np.array([data.user[p[0]] for p in data.training_data], dtype=np.int64),
               train_i=np.array([data.item[p[1]] for p in data.training_data], dtype=np.int64),
               n_item_table=np.array(len(data.item)))
    # one full epoch of the reference sampler (np.random state seeded): every batch, for the sampler property tests
    np.random.seed(seed)
    saved = list(data.training_data)
    batches = [tuple(np.asarray(x, dtype=np.int64) for x in b) for b in next_batch_pairwise(data, bs)]
    data.training_data[:] = saved                         # the sampler shuffles the list in place (util/utils.py:125)
    out["epoch_u"] = np.concatenate([b[0] for b in batches]); out["epoch_i"] = np.concatenate([b[1] for b in batches])
    out["epoch_j"] = np.concatenate([b[2] for b in batches])
    for tag, make in (("lgcn", lambda: LGCN_Encoder(data, D, 3, torch.device("cpu"))), ("mf", lambda: Matrix_Factorization(data, D))):
        torch.manual_seed(seed)
        model = make()
        model.train()
        opt = torch.optim.Adam(model.parameters(), lr=lr)
        out[f"{tag}_E0_user"] = model.embedding_dict["user_emb"].detach().numpy().copy()
        out[f"{tag}_E0_item"] = model.embedding_dict["item_emb"].detach().numpy().copy()
        for s in range(n_steps):
            user_idx, pos_idx, neg_idx = (b.tolist() for b in batches[s])
            rec_user_emb, rec_item_emb = model()
            user_emb, pos_item_emb, neg_item_emb = rec_user_emb[user_idx], rec_item_emb[pos_idx], rec_item_emb[neg_idx]
            bl, rl = bpr_loss(user_emb, pos_item_emb, neg_item_emb), l2_reg_loss(reg, user_emb, pos_item_emb, neg_item_emb)
            batch_loss = bl + rl
            opt.zero_grad()
            batch_loss.backward()
            opt.step()
            out[f"{tag}_loss{s}"] = np.array([batch_loss.item(), bl.item(), rl.item()], dtype=np.float64)
            out[f"{tag}_grad{s}_user"] = model.embedding_dict["user_emb"].grad.numpy().copy()
            out[f"{tag}_grad{s}_item"] = model.embedding_dict["item_emb"].grad.numpy().copy()
            out[f"{tag}_param{s}_user"] = model.embedding_dict["user_emb"].detach().numpy().copy()
            out[f"{tag}_param{s}_item"] = model.embedding_dict["item_emb"].detach().numpy().copy()
        for name, key in (("user", "user_emb"), ("item", "item_emb")):
            st = opt.state[model.embedding_dict[key]]
            out[f"{tag}_exp_avg_{name}"], out[f"{tag}_exp_avg_sq_{name}"] = st["exp_avg"].numpy().copy(), st["exp_avg_sq"].numpy().copy()
    for s in range(n_steps):
        out[f"batch{s}_u"], out[f"batch{s}_i"], out[f"batch{s}_j"] = batches[s]
    return out


def _randomise_bn(module, g):
    for m in module.modules():
        if isinstance(m, torch.nn.BatchNorm1d):
            m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.05)
            m.running_var.copy_(torch.rand(m.num_features, generator=g) * 0.5 + 0.5)
            m.weight.data.copy_(torch.rand(m.num_features, generator=g) + 0.5)
            m.bias.data.copy_(torch.randn(m.num_features, generator=g) * 0.1)


def _widen(module, g, std=0.15):
    """Reference initialisers give ~1e-2 weights (outputs ~1e-3); rescale so tanh/BN are exercised."""
    for m in module.modules():
        if isinstance(m, torch.nn.Linear):
            m.weight.data.copy_(torch.randn(m.weight.shape, generator=g) * std)
            m.bias.data.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)


def golden_towers(work, name, data, user_emb, item_emb, seed):
    from model.DropoutNet import DropoutNet_Learner
    from model.Heater import Heater_Learner
    from model.GAR import GAR_Learner
    from model.ALDI import ALDI_Learner
    os.makedirs(os.path.join(work, "emb"), exist_ok=True)
    torch.save(torch.nn.Parameter(user_emb.clone()), os.path.join(work, "emb", f"{name}_cold_item_MF_user_emb.pt"))
    torch.save(torch.nn.Parameter(item_emb.clone()), os.path.join(work, "emb", f"{name}_cold_item_MF_item_emb.pt"))
    cwd = os.getcwd(); os.chdir(work)
    g = torch.Generator().manual_seed(seed)
    dev = torch.device("cpu")
    out = dict(item_content=np.asarray(data.mapped_item_content, dtype=np.float32), user_emb=user_emb.numpy(), item_emb=item_emb.numpy(),
               cold_idx=np.asarray(data.mapped_cold_item_idx, dtype=np.int64))
    try:
        _orig_load = torch.load
        torch.load = lambda *a, **k: _orig_load(*a, **{**k, "weights_only": False})
        with torch.no_grad():
            args = make_args(name, "item", model="DropoutNet", n_dropout=0.5)
            m = DropoutNet_Learner(args, data, D, dev).eval()
            _widen(m.deepcf_encoder, g); _randomise_bn(m, g)
            for k, v in m.deepcf_encoder.state_dict().items():
                out[f"dn.{k}"] = v.numpy().copy()
            u, i = m.forward()
            out["dn_user_out"], out["dn_item_out"] = u.numpy().copy(), i.numpy().copy()

            args = make_args(name, "item", model="Heater", n_dropout=0.5, alpha=1e-4, n_expert=5)
            m = Heater_Learner(args, data, D, dev).eval()
            _widen(m.heater_encoder, g)
            for k, v in m.heater_encoder.state_dict().items():
                out[f"ht.{k}"] = v.numpy().copy()
            out["ht_n_expert"], out["ht_n_dropout"] = np.array(5), np.array(0.5)
            u, i = m.forward()
            out["ht_user_out"], out["ht_item_out"] = u.numpy().copy(), i.numpy().copy()

            args = make_args(name, "item", model="GAR", alpha=0.05, beta=0.1)
            m = GAR_Learner(args, data, D, dev).eval()
            _widen(m.generator, g)
            for k, v in m.generator.state_dict().items():
                out[f"gar.{k}"] = v.numpy().copy()
            out["gar_cold_out"] = m.generate_item_emb(data.mapped_cold_item_idx).numpy().copy()

            args = make_args(name, "item", model="ALDI", freq_coef_M=4.0, tws=0)
            m = ALDI_Learner(args, data, D, dev).eval()
            _widen(m.user_tower, g); _widen(m.item_tower, g); _randomise_bn(m, g)
            for k, v in m.user_tower.state_dict().items():
                out[f"aldi_u.{k}"] = v.numpy().copy()
            for k, v in m.item_tower.state_dict().items():
                out[f"aldi_i.{k}"] = v.numpy().copy()
            out["aldi_user_out"] = m.get_generated_user_embs().numpy().copy()
            out["aldi_cold_item_out"] = m.get_generated_item_embs(data.mapped_cold_item_idx).numpy().copy()
    finally:
        torch.load = _orig_load
        os.chdir(cwd)
    return out


def golden_cgrc(seed):
    """CGRC's and FSGNN's propagation variants (model/CGRC.py:64-93, model/FSGNN.py:433-442) on the adjacency and the
    embeddings of the committed graph.npz: the layer list with cold item rows frozen at their content vector, the mean
    over layers 0..L, and FSGNN's stack(dim=0).mean — all by the reference's own functions."""
    import scipy.sparse as sp
    import importlib
    ref_cgrc = importlib.import_module("model.CGRC")         # `from model import CGRC` is the trainer class
    FSGNN_Learner = importlib.import_module("model.FSGNN").FSGNN_Learner
    g = dict(np.load(os.path.join(OUT, "graph.npz")))
    n_u, n_i = int(g["user_num"]), int(g["item_num"])
    adj = sp.csr_matrix((g["adj_data"], g["adj_indices"], g["adj_indptr"]), shape=(n_u + n_i, n_u + n_i))
    adj_t = ref_cgrc._sparse_adj_tensor(adj, "cpu")
    rng = np.random.default_rng(seed)
    cold = np.sort(rng.choice(n_i, n_i // 5, replace=False)).astype(np.int64)
    item_x = (rng.standard_normal((n_i, D)) * 0.1).astype(np.float32)          # content-projected item vectors x_i
    U = torch.from_numpy(g["E0_user"])
    out = {"cold_item_idx": cold, "item_x": item_x}
    layers = ref_cgrc._propagate_gprime_frozen_cold(adj_t, U, torch.from_numpy(item_x), n_u, 3, torch.from_numpy(cold))
    for k, h in enumerate(layers):
        out[f"frozen_L{k}"] = h.numpy().copy()
    layers0 = ref_cgrc._propagate_gprime_frozen_cold(adj_t, U, torch.from_numpy(item_x), n_u, 2, torch.zeros(0, dtype=torch.long))
    out["frozen_nocold_L2"] = layers0[-1].numpy().copy()
    zu, zi = ref_cgrc._lightgcn_mean_all_layers(adj_t, U, torch.from_numpy(item_x), n_u, 3)
    out["mean_user"], out["mean_item"] = zu.numpy().copy(), zi.numpy().copy()
    fake = types.SimpleNamespace(n_layers=2, adj_complete=adj_t)
    fu, fi = FSGNN_Learner._lightgcn(fake, U, torch.from_numpy(item_x))
    out["fsgnn_user"], out["fsgnn_item"] = fu.numpy().copy(), fi.numpy().copy()
    return out


def golden_disk(work):
    """The reference's ON-DISK layout for a tiny dataset, written by its own scripts (data/split.py, data/convert.py) from the
    same seeded interactions as eval_tiny.npz: data/<ds>/cold_item/{warm_train,warm_val,...}.csv + info_dict.pkl and
    data/<ds>/<ds>_item_content.npy (main.py:28-52), plus what the reference's loaders + ColdStartDataBuilder make of it."""
    import shutil
    name, cold_object = "syntiny", "item"
    rng = np.random.default_rng(31)
    prepare_dataset(work, name, cold_object, rng, n_users=40, n_items=36, n_inter=420, content_dim=8)
    data, splits, info, content = load_config(work, name, cold_object)
    dst = os.path.join(OUT, "disk", "data", name)
    shutil.rmtree(os.path.join(OUT, "disk"), ignore_errors=True)
    shutil.copytree(os.path.join(work, "data", name, f"cold_{cold_object}"), os.path.join(dst, f"cold_{cold_object}"))
    shutil.copy(os.path.join(work, "data", name, f"{name}_{cold_object}_content.npy"), dst)
    expect = dict(id2user=np.array([data.id2user[i] for i in range(len(data.user))]),
                  id2item=np.array([data.id2item[i] for i in range(len(data.item))]),
                  mapped_warm_item_idx=np.asarray(data.mapped_warm_item_idx), mapped_cold_item_idx=np.asarray(data.mapped_cold_item_idx),
                  mapped_warm_user_idx=np.asarray(data.mapped_warm_user_idx), mapped_cold_user_idx=np.asarray(data.mapped_cold_user_idx),
                  mapped_item_content=np.asarray(data.mapped_item_content)[:len(data.item)],
                  norm_adj_indptr=data.norm_adj.tocsr().indptr, norm_adj_indices=data.norm_adj.tocsr().indices,
                  norm_adj_data=data.norm_adj.tocsr().data, user_num=np.array(data.user_num), item_num=np.array(data.item_num),
                  n_train=np.array(len(data.training_data)))
    np.savez_compressed(os.path.join(OUT, "disk", "expect.npz"), **expect)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None, choices=[None, "train", "cgrc", "disk"], help="regenerate a single fixture")
    only = ap.parse_args().only
    _import_reference()
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)           # fixtures must not depend on thread-count-dependent blocking
    if only == "cgrc":
        np.savez_compressed(os.path.join(OUT, "cgrc.npz"), **golden_cgrc(15))
        return
    if only == "disk":
        with tempfile.TemporaryDirectory() as work:
            golden_disk(work)
        return
    with tempfile.TemporaryDirectory() as work:
        ev, data = golden_eval(work, "synitem", "item", 11, n_users=150, n_items=240, n_inter=4200, content_dim=24)
        np.savez_compressed(os.path.join(OUT, "train.npz"), **golden_train(data, 14))
        if only == "train":
            return
        np.savez_compressed(os.path.join(OUT, "eval_item.npz"), **ev)
        np.savez_compressed(os.path.join(OUT, "graph.npz"), **golden_graph(data, 12))
        np.savez_compressed(os.path.join(OUT, "cgrc.npz"), **golden_cgrc(15))
        tw = golden_towers(work, "synitem", data, torch.from_numpy(ev["user_emb"]), torch.from_numpy(ev["item_emb"]), 13)
        np.savez_compressed(os.path.join(OUT, "towers.npz"), **tw)
        eu, _ = golden_eval(work, "synuser", "user", 21, n_users=160, n_items=130, n_inter=3000, content_dim=16, variants=False)
        np.savez_compressed(os.path.join(OUT, "eval_user.npz"), **eu)
        # fewer unmasked candidates than K in the 'cold' setting: masked ids must appear in the lists
        et, _ = golden_eval(work, "syntiny", "item", 31, n_users=40, n_items=36, n_inter=420, content_dim=8, variants=False)
        np.savez_compressed(os.path.join(OUT, "eval_tiny.npz"), **et)
    with tempfile.TemporaryDirectory() as work:
        golden_disk(work)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
