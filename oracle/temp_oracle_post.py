"""CPU oracle (test infrastructure, like oracle/temp_oracle.py) for the post-ensemble / impute variants of the GRU
families -- SURVEY.md section 8 rows a7, a9, (b), (f)4.

Restates, in torch-CPU fp32 and dense-history faithful:
  encoder calls   models/RRGCN.py:219-272 (forward_post_ensemble, forward_post_ensemble_isolated,
                  forward_isolated_impute, calc_impute_weight) with the layer returns of RRGCN.py:77-116;
                  models/BiRRGCN.py:259-338 (+ forward_post_ensemble_one_direction) with BiRRGCN.py:27-100
  drivers         models/PostDynamicRGCN.py:20-143 (ImputeDynamicRGCN), 146-462 (PostDynamicRGCN /
                  PostEnsembleDynamicRGCN: frequency-gated ensemble of the "local" and the recurrent stream);
                  models/PostBiDynamicRGCN.py:23-372 (the Bi twins)
  frequency stats utils/DropEdge.py:34-82 + utils/frequency.py:31-54 (counts per timestamp, aggregated over the window)
  ranking         utils/post_evaluation.py:77-134 (PostEnsembleEvaluationFilter)
Pinned to outputs of the unmodified reference run on the API stubs (tests/golden/post_*.npz, impute_*.npz).

Only the GRU flavours exist upstream in this mode (the linear RRGCNLayer returns two values where the post-ensemble
callers unpack three).
"""
from __future__ import annotations

from collections import defaultdict
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import temp_oracle as orc
from .temp_oracle import (OracleModel, batch_graphs, decay_state, gru_step, rgcn_layer_graph, rgcn_layer_isolated, time_rows,
                          window_backward, window_forward)

Tensor = torch.Tensor


def post_param_shapes(cfg, impute: bool, post_ensemble: bool) -> Dict[str, tuple]:
    shp = orc.param_shapes(cfg)
    if impute:                                               # RRGCN.py:189-190, BiRRGCN.py:205-208
        names = ["impute_weight"] if not cfg.bidirectional else ["impute_weight_forward", "impute_weight_backward"]
        for n in names:
            shp["ent_encoder.%s.weight" % n] = (1, 1)
            shp["ent_encoder.%s.bias" % n] = (1,)
    if post_ensemble:                                        # PostDynamicRGCN.py:328-338
        for n in ("subject_linear", "object_linear"):
            shp[n + ".0.weight"], shp[n + ".0.bias"] = (3, 3), (3,)
            shp[n + ".2.weight"], shp[n + ".2.bias"] = (1, 3), (1,)
    return shp


class FrequencyStats(object):
    """utils/DropEdge.py:34-82: per target timestamp, how often a subject / object / relation / (subject, relation) /
    (object, relation) of THAT timestamp's training facts occurs in the other timestamps of its window."""

    def __init__(self, quads: np.ndarray, seq_len: int, future: bool):
        per_time = {k: defaultdict(lambda: defaultdict(int)) for k in ("sub", "obj", "rel", "sub_rel", "obj_rel")}
        times = sorted(set(int(q[3]) for q in quads))
        for s, r, o, t in quads.tolist():
            per_time["sub"][t][s] += 1
            per_time["obj"][t][o] += 1
            per_time["rel"][t][r] += 1
            per_time["sub_rel"][t][(s, r)] += 1
            per_time["obj_rel"][t][(o, r)] += 1
        self.agg = {k: defaultdict(lambda: defaultdict(int)) for k in per_time}
        max_step = len(times)
        for tt in times:
            upper = tt if not future else min(max_step + 1, tt + seq_len)
            for cur in range(max(0, tt - seq_len + 1), upper):
                if cur == tt:
                    continue
                for k, table in per_time.items():
                    here = table[cur]
                    for item in list(table[tt].keys()):
                        if item in here:
                            self.agg[k][tt][item] += here[item]

    def features(self, triples_global: Sequence, t: int):
        """-> (sub_features [n, 3], obj_features [n, 3]) of PostDynamicRGCN.py:430-448."""
        a = self.agg
        sub, obj = [], []
        for s, r, o in triples_global:
            sub.append([a["obj"][t][o], a["rel"][t][r], a["obj_rel"][t][(o, r)]])
            obj.append([a["sub"][t][s], a["rel"][t][r], a["sub_rel"][t][(s, r)]])
        return torch.tensor(sub, dtype=torch.float32).view(-1, 3), torch.tensor(obj, dtype=torch.float32).view(-1, 3)


class PostOracle(OracleModel):
    def __init__(self, cfg, params, graph_dict_train, impute: bool, post_ensemble: bool, train_quads: Optional[np.ndarray] = None):
        super().__init__(cfg, params, graph_dict_train)
        assert cfg.gru, "post-ensemble / impute exist for the GRU flavours only"
        self.impute, self.post_ensemble = bool(impute), bool(post_ensemble)
        self.freq = FrequencyStats(train_quads, cfg.seq_len, cfg.bidirectional) if post_ensemble else None

    # ---- encoder calls ------------------------------------------------------------------------------------------
    def impute_weights(self, dts: List[Tensor]) -> List[Tensor]:
        p = self.p
        if not self.cfg.bidirectional:                       # RRGCN.py:271-272
            w, b = p["ent_encoder.impute_weight.weight"], p["ent_encoder.impute_weight.bias"]
            return [torch.exp(-torch.clamp(dts[0] * w.view(()) + b.view(()), min=0))]
        out = []                                             # BiRRGCN.py:309-310, 331-332
        for dt, n in zip(dts, ("impute_weight_forward", "impute_weight_backward")):
            w, b = p["ent_encoder.%s.weight" % n], p["ent_encoder.%s.bias" % n]
            out.append(torch.exp(-torch.clamp(dt * w.view(()) + b.view(()), min=0)) / 2)
        return out

    @staticmethod
    def _blend(ws: List[Tensor], locs: List[Tensor], x: Tensor) -> Tensor:
        rest = 1
        acc = 0
        for w, loc in zip(ws, locs):
            acc = acc + w * loc
            rest = rest - w
        return acc + rest * x

    def enc_post_ensemble(self, graphs, times, prev1, prev2, dts, direction: Optional[str] = "forward"):
        """RRGCN.py:219-234, BiRRGCN.py:259-293 -> (second_local + te, first, second + te); ``first`` aliases ``second``
        (the layer-2 cell writes into the graph object layer 1 returned)."""
        cfg, p = self.cfg, self.p
        bg = batch_graphs(graphs)
        sizes = [g.num_nodes for g in graphs]
        names = self._rnn_names(direction)
        h0 = self._embed(bg)
        if cfg.rec_only_last_layer:
            first = rgcn_layer_graph(p, "ent_encoder.layer_1.", cfg, h0, bg, relu=False)
        else:
            _, first = self._rec_layer_graph("ent_encoder.layer_1.", h0, bg, prev1, dts, names, relu=False)
            if cfg.use_time_embedding:
                first = first + time_rows(p, "ent_encoder.layer_1.", times, sizes)
        local, second = self._rec_layer_graph("ent_encoder.layer_2.", first, bg, prev2, dts, names, relu=cfg.layer2_relu)
        if cfg.use_time_embedding:
            te = time_rows(p, "ent_encoder.layer_2.", times, sizes)
            local, second = local + te, second + te
        return local, second, second

    def _iso_first(self, t, prev1, dts, names):
        cfg, p = self.cfg, self.p
        x = p["ent_embeds"]
        if cfg.rec_only_last_layer:
            return rgcn_layer_isolated(p, "ent_encoder.layer_1.", cfg, x, relu=False)
        pre = "ent_encoder.layer_1."
        x1 = rgcn_layer_isolated(p, pre, cfg, x, relu=False)
        first = 0
        for prev, dt, rn in zip(prev1, dts, names):
            first = first + gru_step(p, pre + rn + ".", cfg, x1, decay_state(p, pre, cfg, prev, dt))
        return first + p[pre + "time_embed"][int(t)] if cfg.use_time_embedding else first

    def enc_post_ensemble_isolated(self, t, prev1, prev2, dts, locs):
        """RRGCN.py:236-254, BiRRGCN.py:295-319 -> (second_local (imputed when the encoder has impute) + te, second + te)."""
        cfg, p = self.cfg, self.p
        names = self._rnn_names(None if cfg.bidirectional else "forward")
        pre = "ent_encoder.layer_2."
        first = self._iso_first(t, prev1, dts, names)
        local = rgcn_layer_isolated(p, pre, cfg, first, relu=cfg.layer2_relu)
        rec = 0
        for prev, dt, rn in zip(prev2, dts, names):
            rec = rec + gru_step(p, pre + rn + ".", cfg, local, decay_state(p, pre, cfg, prev, dt))
        if self.impute:
            local = self._blend(self.impute_weights(dts), locs, local)
        if cfg.use_time_embedding:
            te = p[pre + "time_embed"][int(t)]
            local, rec = local + te, rec + te
        return local, rec

    def enc_isolated_impute(self, t, prev1, prev2, dts, locs):
        """RRGCN.py:256-269 + 105-116, BiRRGCN.py:321-338 + 83-100: the layer-2 cells read the IMPUTED local stream."""
        cfg, p = self.cfg, self.p
        names = self._rnn_names(None if cfg.bidirectional else "forward")
        pre = "ent_encoder.layer_2."
        first = self._iso_first(t, prev1, dts, names)
        x = rgcn_layer_isolated(p, pre, cfg, first, relu=cfg.layer2_relu)
        x = self._blend(self.impute_weights(dts), locs, x)
        rec = 0
        for prev, dt, rn in zip(prev2, dts, names):
            rec = rec + gru_step(p, pre + rn + ".", cfg, x, decay_state(p, pre, cfg, prev, dt))
        return rec + p[pre + "time_embed"][int(t)] if cfg.use_time_embedding else rec

    # ---- drivers ------------------------------------------------------------------------------------------------
    def _scan_post(self, time_batched, direction: str):
        """PostDynamicRGCN.py:60-79 / PostBiDynamicRGCN.py:73-98 (eval: full graphs)."""
        cfg = self.cfg
        L, bsz = cfg.seq_len, len(time_batched[0])
        loc = torch.zeros(bsz, self.M, self.D)
        rec = torch.zeros(bsz, 2, self.M, self.D)
        start = torch.zeros(bsz, self.M)
        for k in range(L - 1):
            ts = [t for t in time_batched[k] if t is not None]
            if not ts:
                continue
            graphs = [self.gd[t] for t in ts]
            p1, p2, dt = self._gather_prev(graphs, rec, start, k)
            local, first, second = self.enc_post_ensemble(graphs, ts, [p1], [p2], [dt], direction)
            loc = torch.zeros(bsz, self.M, self.D)           # PostDynamicRGCN.py:34-35: fresh zeros every step
            rec = torch.zeros(bsz, 2, self.M, self.D)
            off = 0
            for i, g in enumerate(graphs):
                idx = torch.from_numpy(g.ids)
                loc[i][idx] = local[off:off + g.num_nodes]
                rec[i][0][idx] = first[off:off + g.num_nodes]
                rec[i][1][idx] = second[off:off + g.num_nodes]
                start[i][idx] = k
                off += g.num_nodes
        if direction == "backward":
            loc, rec, start = torch.flip(loc, [0]), torch.flip(rec, [0]), torch.flip(start, [0])
        return loc, rec, start

    def evaluate_embed_post(self, t_list: Sequence[int]):
        cfg = self.cfg
        L = cfg.seq_len
        tb_f = window_forward(t_list, L, self.times)
        ts = tb_f[-1]
        graphs = [self.gd[t] for t in ts]
        sizes = [g.num_nodes for g in graphs]
        res = {"times": ts, "graphs": graphs}
        loc_f, rec_f, start_f = self._scan_post(tb_f, "forward")
        res.update(loc_f=loc_f, rec_f=rec_f, start_f=start_f)
        f1, f2, dtf = self._gather_prev(graphs, rec_f, start_f, L - 1)
        if cfg.bidirectional:
            loc_b, rec_b, start_b = self._scan_post(window_backward(t_list, L, self.times), "backward")
            res.update(loc_b=loc_b, rec_b=rec_b, start_b=start_b)
            b1, b2, dtb = self._gather_prev(graphs, rec_b, start_b, L - 1)
            local, _, second = self.enc_post_ensemble(graphs, ts, [f1, b1], [f2, b2], [dtf, dtb], direction=None)
        else:
            local, _, second = self.enc_post_ensemble(graphs, ts, [f1], [f2], [dtf], "forward")
        res["per_graph_loc"], res["per_graph"] = list(local.split(sizes)), list(second.split(sizes))
        return res

    def _iso_args(self, res, i):
        L = self.cfg.seq_len
        dirs = ["f", "b"] if self.cfg.bidirectional else ["f"]
        prev1 = [res["rec_" + d][i][0] for d in dirs]
        prev2 = [res["rec_" + d][i][1] for d in dirs]
        dts = [(L - 1 - res["start_" + d][i]).unsqueeze(-1) for d in dirs]
        locs = [res["loc_" + d][i] for d in dirs]
        return prev1, prev2, dts, locs

    def all_embeds_impute(self, res, i: int) -> Tensor:
        """ImputeDynamicRGCN.get_all_embeds_Gt (PostDynamicRGCN.py:24-31, PostBiDynamicRGCN.py:30-40)."""
        g, t = res["graphs"][i], res["times"][i]
        out = self.enc_isolated_impute(t, *self._iso_args(res, i)).clone()
        out[torch.from_numpy(g.ids)] = res["per_graph"][i]
        return out

    def all_embeds_post(self, res, i: int):
        """PostDynamicRGCN.get_all_embeds_Gt (PostDynamicRGCN.py:174-187, PostBiDynamicRGCN.py:181-191) -> (local, recurrent)."""
        g, t = res["graphs"][i], res["times"][i]
        loc, rec = self.enc_post_ensemble_isolated(t, *self._iso_args(res, i))
        loc, rec = loc.clone(), rec.clone()
        idx = torch.from_numpy(g.ids)
        loc[idx] = res["per_graph_loc"][i]
        rec[idx] = res["per_graph"][i]
        return loc, rec

    def ensemble_weights(self, samples: Tensor, g, t: int):
        """PostDynamicRGCN.py:430-462: sigmoid of a 3 -> 3 -> 1 MLP over the frequency features."""
        p = self.p
        tri = [(int(g.ids[s]), int(r), int(g.ids[o])) for s, r, o in samples.tolist()]
        sub_f, obj_f = self.freq.features(tri, int(t))

        def mlp(x, name):
            h = torch.relu(x @ p[name + ".0.weight"].t() + p[name + ".0.bias"])
            return torch.sigmoid(h @ p[name + ".2.weight"].t() + p[name + ".2.bias"])
        return mlp(sub_f, "subject_linear"), mlp(obj_f, "object_linear")

    def evaluate_post(self, t_list, valid, test=None, val: bool = True):
        """evaluate(t_list, val) of ImputeDynamicRGCN (PostDynamicRGCN.py:115-143: ordinary filtered ranks on the recurrent
        stream with the imputed all-entity table) or of PostEnsembleDynamicRGCN (367-411: ranks of the weighted sum of the
        masked local and recurrent scores, utils/post_evaluation.py:77-134).  The history index of calc_metrics only
        advances for graphs with edges (as in the plain models)."""
        res = self.evaluate_embed_post(t_list)
        graph_dict = valid if val else test
        rel = self.p["rel_embeds"]
        ranks, losses = [], []
        i = 0
        for pos, t in enumerate(res["times"]):
            g = graph_dict[int(t)]
            shifted = res
            if i != pos:                                     # lagging history index
                shifted = dict(res)
                for key in list(res.keys()):
                    if key.startswith(("loc_", "rec_", "start_")):
                        v = res[key].clone()
                        v[pos] = res[key][i]
                        shifted[key] = v
            if g.num_edges == 0:
                continue
            samples = torch.from_numpy(np.stack([g.src, g.rel, g.dst], axis=1)).long()
            parts = [x for x in (self.gd.get(int(t)), valid.get(int(t)), test.get(int(t)) if test is not None else None)
                     if x is not None]
            ent = res["per_graph"][pos]
            if not self.post_ensemble:
                alls = self.all_embeds_impute(shifted, pos)
                ranks.append(orc.filtered_ranks(orc.score_complex, ent, rel, alls, samples, g, parts))
                s, r, o = ent[samples[:, 0]], rel[samples[:, 1]], ent[samples[:, 2]]
                losses.append(float(torch.nn.functional.binary_cross_entropy_with_logits(orc.score_complex(s, r, o),
                                                                                         torch.ones(samples.shape[0]))))
            else:
                loc_all, rec_all = self.all_embeds_post(shifted, pos)
                w_sub, w_obj = self.ensemble_weights(samples, g, int(t))
                ranks.append(ensemble_ranks(res["per_graph_loc"][pos], ent, rel, loc_all, rec_all, w_sub, w_obj, samples, g, parts))
            i += 1
        ranks = torch.cat(ranks) if ranks else torch.zeros(0, dtype=torch.long)
        return ranks, (float(np.mean(losses)) if losses else float("nan"))


def ensemble_ranks(ent_loc, ent_rec, rel, all_loc, all_rec, w_sub, w_obj, samples, g, splits) -> Tensor:
    """utils/post_evaluation.py:82-134: per side, masked local and recurrent scores (batches of 100 concatenated),
    ``w * local + (1 - w) * recurrent`` -- the OBJECT side (mode 'tail') is weighted with weight_subject, the subject side
    with weight_object (lines 91-94) --, sigmoid, descending sort, position of the target; subject side first, + 1."""
    tri = np.concatenate([np.stack([x.src, x.rel, x.dst], axis=1) for x in splits], axis=0)
    th, tt = {}, {}
    for h, r, t in tri.tolist():
        tt.setdefault((h, r), []).append(t)
        th.setdefault((r, t), []).append(h)
    ids = g.ids
    M, n = all_loc.shape[0], samples.shape[0]
    gids = torch.from_numpy(ids)
    out = {}
    for mode, w in (("tail", w_sub), ("head", w_obj)):
        mask = torch.zeros(n, M, dtype=torch.bool)
        for q, (h, r, t) in enumerate(samples.tolist()):
            if mode == "tail":
                mask[q, torch.from_numpy(ids[np.array(list(set(tt[(h, r)])))])] = True
                mask[q, int(ids[t])] = False
            else:
                mask[q, torch.from_numpy(ids[np.array(list(set(th[(r, t)])))])] = True
                mask[q, int(ids[h])] = False
        r_ = rel[samples[:, 1]]
        scs = []
        for ent, alls in ((ent_loc, all_loc), (ent_rec, all_rec)):
            if mode == "tail":
                sc = orc.score_complex(ent[samples[:, 0]], r_, alls, mode="tail")
            else:
                sc = orc.score_complex(alls, r_, ent[samples[:, 2]], mode="head")
            scs.append(torch.where(mask, -10e6 * torch.ones_like(sc), sc))
        target = gids[samples[:, 2 if mode == "tail" else 0]]
        sc = torch.sigmoid(w * scs[0] + (1 - w) * scs[1])
        _, order = torch.sort(sc, dim=1, descending=True)
        out[mode] = torch.nonzero(order == target.view(-1, 1))[:, 1].view(-1)
    return torch.cat([out["head"], out["tail"]]) + 1
