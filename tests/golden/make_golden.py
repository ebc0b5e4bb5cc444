"""Generates tests/golden/*.npz by running the UNMODIFIED reference (JiapengWu/TeMP) here.

The reference's python files are imported from /root/reference on top of oracle/_stubs (a
restatement of the DGL 0.4.1 / pytorch-lightning 0.5.2 API slices it calls -- those packages are
not installable in this container).  Model parameters are overwritten with oracle.fill_values (an
exact-integer hash), so the golden files hold OUTPUTS only and any implementation can rebuild the
inputs from (case config, dataset files under tests/golden/data).

Run from the repo root (needs /root/reference; the committed .npz files do not):
    python tests/golden/make_golden.py
"""
import os
import sys
import warnings
from argparse import Namespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from tests.golden.cases import (CASES, CPU_CASES, CPU_TRAIN_CASES, POST_CASES, RANK_CASES, SAMPLER_CASES, TRAIN_CASES,  # noqa: E402
                                dataset_path)
from oracle.temp_oracle import fill_values  # noqa: E402


def reference_on_path():
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_stubs"))
    sys.path.insert(0, "/root/reference")


def ref_args(case):
    return Namespace(
        dataset_dir="interpolation", dataset=dataset_path(case["dataset"]), score_function="complex",
        module=case["module"], n_gpu=-1, use_cuda=False, hidden_size=case["D"], embed_size=case["D"],
        dropout=0.1, num_layers=1, lr=1e-3, n_bases=case["n_bases"], rgcn_layers=2,
        train_seq_len=case["L"], test_seq_len=case["L"], batch_size=len(case["t_list"]), seed=123,
        negative_rate=case.get("negative_rate", 5), num_pos_facts=case.get("num_pos_facts", 3000),
        debug=False, rec_only_last_layer=case["rec_only_last_layer"], use_time_embedding=case["use_time_embedding"],
        inv_temperature=0.1, use_embed_for_non_active=case.get("use_embed_for_non_active", False), edge_dropout=False, random_dropout=False,
        type1=case.get("type1", False), post_ensemble=False, post_aggregation=False,
        learnable_lambda=case.get("learnable_lambda", False), impute=False, EMA=False,
        rate_lower=0.2, rate_upper=0.8, lambda_1=2, lambda_2=10, lambda_3=20)


def ref_graph_dicts(args):
    """utils/dataset.py:268-290 of the reference minus the pickle cache (dataset dir is read-only
    for the real files and we do not want cache files in the fixtures)."""
    from utils.dataset import (get_total_number, get_train_val_test_graph_at_t, load_quadruples,
                               load_quadruples_interpolation)
    _, total_times = load_quadruples(args.dataset, "train.txt", "valid.txt", "test.txt")
    t2t = load_quadruples_interpolation(args.dataset, "train.txt", "valid.txt", "test.txt", total_times)
    num_e, num_r = get_total_number(args.dataset, "stat.txt")
    gtr, gva, gte = {}, {}, {}
    for tim in total_times:
        gtr[tim], gva[tim], gte[tim] = get_train_val_test_graph_at_t(t2t[tim], num_r)
    return num_e, num_r, gtr, gva, gte


def run_train_case(tc):
    """loss = model.forward(t_list) of the unmodified reference in train mode, dropout p = 0, fixed global seeds."""
    base = dict(next(c for c in CASES if c["name"] == tc["base"]))
    base.update({k: v for k, v in tc.items() if k in ("negative_rate", "num_pos_facts")})
    model, _ = ref_model(base, dropout=0.0, random_dropout=tc["random_dropout"])
    model.train()
    np.random.seed(tc["seed"])
    torch.manual_seed(tc["seed"])
    with torch.no_grad():
        loss = model.forward(torch.tensor(base["t_list"], dtype=torch.long))
    return {"loss": np.asarray(float(loss), dtype=np.float64)}


def ref_model(case, dropout=None, random_dropout=False):
    from baselines.StaticRGCN import StaticRGCN
    from models.BiDynamicRGCN import BiDynamicRGCN
    from models.BiSelfAttentionRGCN import BiSelfAttentionRGCN
    from models.DynamicRGCN import DynamicRGCN
    from models.SelfAttentionRGCN import SelfAttentionRGCN
    cls = {"SRGCN": StaticRGCN, "GRRGCN": DynamicRGCN, "RRGCN": DynamicRGCN, "BiGRRGCN": BiDynamicRGCN,
           "BiRRGCN": BiDynamicRGCN, "SARGCN": SelfAttentionRGCN, "BiSARGCN": BiSelfAttentionRGCN}[case["module"]]
    args = ref_args(case)
    if dropout is not None:
        args.dropout = dropout
    args.random_dropout = random_dropout
    num_e, num_r, gtr, gva, gte = ref_graph_dicts(args)
    torch.manual_seed(123)
    with torch.no_grad():
        model = cls(args, num_e, num_r, gtr, gva, gte)
        for name, prm in model.state_dict().items():
            prm.copy_(torch.from_numpy(fill_values(name, tuple(prm.shape))))
    model.eval()
    return model, (gtr, gva, gte)


def run_case(case):
    model, _ = ref_model(case)
    t_list = torch.tensor(case["t_list"], dtype=torch.long)
    L = case["L"]
    mod = case["module"]
    out = {}
    with torch.no_grad():
        if mod == "SRGCN":
            per_graph, g_list = model.evaluate_embed(t_list, val=True)
            times = [int(t) for t in case["t_list"]]
            alls = [model.get_all_embeds_Gt(t_list[i], g_list[i], per_graph[i]) for i in range(len(times))]
        elif mod in ("GRRGCN", "RRGCN"):
            per_graph, test_graphs, time_list, hist, start = model.evaluate_embed(t_list, val=True)
            times = [int(t) for t in time_list[-1]]
            alls = [model.get_all_embeds_Gt(per_graph[i], test_graphs[i], time_list[-1][i], hist[i][0], hist[i][1],
                                            L - 1 - start[i]) for i in range(len(times))]
            out["start"] = start.numpy()
            out["hist_abs_sum"] = hist.abs().sum(dim=(2, 3)).numpy()
            out["hist_rows_nonzero"] = (hist.abs().sum(-1) != 0).sum(-1).numpy()
        elif mod in ("BiGRRGCN", "BiRRGCN"):
            per_graph, test_graphs, tl, hf, sf, hb, sb = model.evaluate_embed(t_list, val=True)
            times = [int(t) for t in tl]
            alls = [model.get_all_embeds_Gt(per_graph[i], test_graphs[i], tl[i], hf[i][0], hf[i][1], L - 1 - sf[i],
                                            hb[i][0], hb[i][1], L - 1 - sb[i]) for i in range(len(times))]
            out["start_f"], out["start_b"] = sf.numpy(), sb.numpy()
        else:                                    # SARGCN / BiSARGCN: evaluate() inlined up to calc_metrics
            from models.BiDynamicRGCN import BiDynamicRGCN
            if mod == "SARGCN":
                g_tr, time_list = model.get_batch_graph_list(t_list, L, model.graph_dict_train)
                hist, mask = model.pre_forward(g_tr, time_list, val=True)
                train_graphs, tl = g_tr[-1], time_list[-1]
            else:
                gf, tf, gb, tb = BiDynamicRGCN.get_batch_graph_list(t_list, L, model.graph_dict_train)
                hfw, mfw = model.pre_forward(gf, tf, forward=True)
                hbw, mbw = model.pre_forward(gb, tb, forward=False)
                train_graphs, tl = gf[-1], tf[-1]
                hist = torch.cat([hfw, hbw], dim=0)
                mask = torch.cat([mfw, mbw, mfw.new_zeros(1, *mfw.shape[1:])], dim=0)
            sizes = [len(g.nodes()) for g in train_graphs]
            if mod == "SARGCN":
                per_graph = model.get_final_graph_embeds(train_graphs, tl, sizes, hist, mask, full=True, val=True)
            else:
                per_graph = model.get_final_graph_embeds(train_graphs, tl, sizes, hist, mask, full=True)
            times = [int(t) for t in tl]
            alls = [model.get_all_embeds_Gt(per_graph[i], train_graphs[i], tl[i], hist[:, i, 0], hist[:, i, 1],
                                            mask[:, i], val=True) for i in range(len(times))]
    out["times"] = np.asarray(times, dtype=np.int64)
    out["sizes"] = np.asarray([p.shape[0] for p in per_graph], dtype=np.int64)
    out["per_graph"] = torch.cat(list(per_graph), dim=0).numpy()
    rows = np.asarray(case_rows(case, alls[0].shape[0]), dtype=np.int64)
    out["all_rows"] = rows
    out["all_embeds"] = torch.stack(alls, dim=0)[:, rows].numpy()
    return out


def case_rows(case, M):
    """Row subset of the all-entity table stored in the golden file (all rows when M is small)."""
    if M <= 512:
        return list(range(M))
    return sorted(set(range(0, M, 29)) | set(range(0, 64)))


def run_sampler_case(case):
    """Negative-sample index streams of the reference's CorruptTriples under fixed seeds
    (utils/CorrptTriples.py:26-85) -- the bit-exact contract of SURVEY section 8(a) row N."""
    from utils.CorrptTriples import CorruptTriples
    args = ref_args(dict(case, module="GRRGCN", D=8, n_bases=8, L=2, rec_only_last_layer=True,
                         use_time_embedding=True, t_list=case["times"]))
    num_e, num_r, gtr, _, _ = ref_graph_dicts(args)
    cor = CorruptTriples(args, gtr)
    np.random.seed(case["seed"])
    torch.manual_seed(case["seed"])
    out = {}
    for t in case["times"]:
        tri, nt, nh, lab = cor.single_graph_negative_sampling(torch.tensor(t), gtr[t], num_e)
        out["triples_%d" % t] = tri.numpy()
        out["neg_tail_%d" % t] = nt.numpy()
        out["neg_head_%d" % t] = nh.numpy()
    return out


def run_rank_case(rc):
    """ranks, loss = model.evaluate(t_list, val=True) of the unmodified reference."""
    case = next(c for c in CASES if c["name"] == rc["base"])
    model, (_, gva, _) = ref_model(case)
    if rc.get("empty_first"):
        t_empty = max(int(t) for t in case["t_list"])
        g = gva[t_empty]
        sub = g.edge_subgraph(torch.zeros(0, dtype=torch.long), preserve_nodes=True)
        sub.ids = g.ids
        for k in ("id", "norm"):
            if k in g.ndata:
                sub.ndata[k] = g.ndata[k]
        for k in g.edata:
            sub.edata[k] = g.edata[k][:0]
        model.graph_dict_val = dict(gva)
        model.graph_dict_val[t_empty] = sub
    with torch.no_grad():
        ranks, loss = model.evaluate(torch.tensor(case["t_list"], dtype=torch.long), val=True)
    return {"ranks": ranks.numpy().astype(np.int64), "loss": np.asarray(float(loss), dtype=np.float64)}


def run_post_case(pc):
    """evaluate_embed / get_all_embeds_Gt / evaluate of the unmodified Impute* / PostEnsemble* model classes."""
    from models.PostBiDynamicRGCN import ImputeBiDynamicRGCN, PostEnsembleBiDynamicRGCN
    from models.PostDynamicRGCN import ImputeDynamicRGCN, PostEnsembleDynamicRGCN
    case = next(c for c in CASES if c["name"] == pc["base"])
    bi = case["module"].startswith("Bi")
    cls = ((PostEnsembleBiDynamicRGCN if bi else PostEnsembleDynamicRGCN) if pc["post_ensemble"]
           else (ImputeBiDynamicRGCN if bi else ImputeDynamicRGCN))
    args = ref_args(case)
    args.impute, args.post_ensemble = pc["impute"], pc["post_ensemble"]
    num_e, num_r, gtr, gva, gte = ref_graph_dicts(args)
    torch.manual_seed(123)
    with torch.no_grad():
        model = cls(args, num_e, num_r, gtr, gva, gte)
        for name, prm in model.state_dict().items():
            prm.copy_(torch.from_numpy(fill_values(name, tuple(prm.shape))))
    model.eval()
    t_list = torch.tensor(case["t_list"], dtype=torch.long)
    L = case["L"]
    out = {}
    with torch.no_grad():
        ev = model.evaluate_embed(t_list, val=True)
        if not bi:
            if pc["post_ensemble"]:
                loc, rec, graphs, time_list, hl, hr, st = ev
                tl = time_list[-1]
                alls = [model.get_all_embeds_Gt(loc[i], rec[i], graphs[i], tl[i], hl[i], hr[i][0], hr[i][1], L - 1 - st[i])
                        for i in range(len(tl))]
            else:
                rec, graphs, time_list, hl, hr, st = ev
                loc, tl = None, time_list[-1]
                alls = [model.get_all_embeds_Gt(rec[i], graphs[i], tl[i], hl[i], hr[i][0], hr[i][1], L - 1 - st[i])
                        for i in range(len(tl))]
        else:
            if pc["post_ensemble"]:
                loc, rec, graphs, tl, hfl, hfr, sf, hbl, hbr, sb = ev
                alls = [model.get_all_embeds_Gt(loc[i], rec[i], graphs[i], tl[i], hfl[i], hfr[i][0], hfr[i][1], L - 1 - sf[i],
                                                hbl[i], hbr[i][0], hbr[i][1], L - 1 - sb[i]) for i in range(len(tl))]
            else:
                rec, graphs, tl, hfl, hfr, sf, hbl, hbr, sb = ev
                loc = None
                alls = [model.get_all_embeds_Gt(rec[i], graphs[i], tl[i], hfl[i], hfr[i][0], hfr[i][1], L - 1 - sf[i],
                                                hbl[i], hbr[i][0], hbr[i][1], L - 1 - sb[i]) for i in range(len(tl))]
        ranks, _ = model.evaluate(t_list, val=True)
    out["times"] = np.asarray([int(t) for t in tl], dtype=np.int64)
    out["sizes"] = np.asarray([p.shape[0] for p in rec], dtype=np.int64)
    out["per_graph"] = torch.cat(list(rec), dim=0).numpy()
    if loc is not None:
        out["per_graph_loc"] = torch.cat(list(loc), dim=0).numpy()
    M = alls[0][0].shape[0] if pc["post_ensemble"] else alls[0].shape[0]
    rows = np.asarray(case_rows(case, M), dtype=np.int64)
    out["all_rows"] = rows
    if pc["post_ensemble"]:
        out["all_embeds_loc"] = torch.stack([a[0] for a in alls], dim=0)[:, rows].numpy()
        out["all_embeds"] = torch.stack([a[1] for a in alls], dim=0)[:, rows].numpy()
    else:
        out["all_embeds"] = torch.stack(alls, dim=0)[:, rows].numpy()
    out["ranks"] = ranks.numpy().astype(np.int64)
    return out


def main():
    warnings.filterwarnings("ignore")
    reference_on_path()
    only_train = "--train-only" in sys.argv
    only_rank = "--rank-only" in sys.argv
    if "--only" in sys.argv:                       # --only name1,name2: just these cases (of any kind)
        names = set(sys.argv[sys.argv.index("--only") + 1].split(","))
        for fn, cases in ((run_post_case, POST_CASES), (run_rank_case, RANK_CASES), (run_case, CASES + CPU_CASES), (run_train_case, TRAIN_CASES + CPU_TRAIN_CASES),
                          (run_sampler_case, SAMPLER_CASES)):
            for case in cases:
                if case["name"] in names:
                    res = fn(case)
                    path = os.path.join(HERE, case["name"] + ".npz")
                    np.savez_compressed(path, **res)
                    print("%-40s %.1f KB" % (case["name"], os.path.getsize(path) / 1024))
        return
    for rc in ([] if only_train else RANK_CASES):
        res = run_rank_case(rc)
        path = os.path.join(HERE, rc["name"] + ".npz")
        np.savez_compressed(path, **res)
        print("%-40s ranks=%d loss=%.9g" % (rc["name"], res["ranks"].shape[0], float(res["loss"])))
    if only_rank:
        return
    for case in ([] if only_train else CASES + CPU_CASES):
        res = run_case(case)
        path = os.path.join(HERE, case["name"] + ".npz")
        np.savez_compressed(path, **res)
        print("%-40s rows=%d  %.1f KB" % (case["name"], res["per_graph"].shape[0], os.path.getsize(path) / 1024))
    for case in TRAIN_CASES + CPU_TRAIN_CASES:
        res = run_train_case(case)
        path = os.path.join(HERE, case["name"] + ".npz")
        np.savez_compressed(path, **res)
        print("%-40s loss=%.9g" % (case["name"], float(res["loss"])))
    for case in SAMPLER_CASES:
        res = run_sampler_case(case)
        path = os.path.join(HERE, case["name"] + ".npz")
        np.savez_compressed(path, **res)
        print("%-40s %.1f KB" % (case["name"], os.path.getsize(path) / 1024))
    for case in POST_CASES:
        res = run_post_case(case)
        path = os.path.join(HERE, case["name"] + ".npz")
        np.savez_compressed(path, **res)
        print("%-40s ranks=%d  %.1f KB" % (case["name"], res["ranks"].shape[0], os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
