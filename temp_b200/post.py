"""The post-ensemble / impute model shells of the GRU families, on the CUDA path (SURVEY.md section 8 rows a7, a9, (b),
(f)4; main.py:57-72 selects them with ``--impute`` / ``--post-ensemble``):

    ImputeDynamicRGCN            models/PostDynamicRGCN.py:20-143
    PostEnsembleDynamicRGCN      models/PostDynamicRGCN.py:146-462 (the --post-ensemble class; the frequency-gated ensemble
                                 of the "local" and the recurrent stream)
    ImputeBiDynamicRGCN          models/PostBiDynamicRGCN.py:23-175
    PostEnsembleBiDynamicRGCN    models/PostBiDynamicRGCN.py:178-372

The window forward is the same launch program as the plain models (``encode``); its layer-2 launches additionally store
their output BEFORE the recurrent cells (+ time embedding) -- the "local" stream (models/RRGCN.py:227-233) -- into
``res.bufs['local']``.  The all-entity tables go through the encoder's ``forward_isolated_impute`` /
``forward_post_ensemble_isolated`` calls (temp_b200/stepwise.py) with the reference's dense argument lists, so
``get_all_embeds_Gt`` keeps the reference's signatures.  The frequency features (utils/DropEdge.py:34-82,
utils/frequency.py:31-54) are counted once per model from the training snapshots; the ensemble ranking restates
utils/post_evaluation.py:77-134 in torch on the GPU.  With gradients enabled ``forward`` raises (the autograd fallback of
temp_b200/autograd_path.py covers the plain models); ``--post-aggregation`` (PostDynamicRGCN proper, four MLPs) is not built.
"""
from __future__ import annotations

from collections import defaultdict

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .models import BiDynamicRGCN, DynamicRGCN


class FrequencyStats(object):
    """utils/DropEdge.py:34-82: per target timestamp, how often a subject / object / relation / (subject, relation) /
    (object, relation) of THAT timestamp's training facts occurs in the other timestamps of its window (``future``: the
    window also reaches forward, the Bi models)."""

    def __init__(self, graph_dict_train, seq_len: int, future: bool):
        keys = ("sub", "obj", "rel", "sub_rel", "obj_rel")
        per_time = {k: defaultdict(lambda: defaultdict(int)) for k in keys}
        times = sorted(int(t) for t in graph_dict_train.keys())
        for t in times:
            g = graph_dict_train[t]
            ids = g.node_ids
            for s, r, o in zip(ids[g.src].tolist(), np.asarray(g.rel).tolist(), ids[g.dst].tolist()):
                per_time["sub"][t][s] += 1
                per_time["obj"][t][o] += 1
                per_time["rel"][t][r] += 1
                per_time["sub_rel"][t][(s, r)] += 1
                per_time["obj_rel"][t][(o, r)] += 1
        self.agg = {k: defaultdict(lambda: defaultdict(int)) for k in keys}
        max_step = len(times)
        for tt in times:
            upper = tt if not future else min(max_step + 1, tt + seq_len)
            for cur in range(max(0, tt - seq_len + 1), upper):
                if cur == tt:
                    continue
                for k in keys:
                    here = per_time[k][cur]
                    for item in list(per_time[k][tt].keys()):
                        if item in here:
                            self.agg[k][tt][item] += here[item]

    def features(self, triples_global, t: int):
        """-> (sub_features [n, 3], obj_features [n, 3]) of models/PostDynamicRGCN.py:430-448."""
        a = self.agg
        sub = [[a["obj"][t][o], a["rel"][t][r], a["obj_rel"][t][(o, r)]] for s, r, o in triples_global]
        obj = [[a["sub"][t][s], a["rel"][t][r], a["sub_rel"][t][(s, r)]] for s, r, o in triples_global]
        return torch.tensor(sub, dtype=torch.float32).view(-1, 3), torch.tensor(obj, dtype=torch.float32).view(-1, 3)


def _freq_mlp():
    return nn.Sequential(nn.Linear(3, 3), nn.ReLU(), nn.Linear(3, 1))


class _PostBase(object):
    """Shared pieces of the four shells (mixed in before the plain shell)."""
    post_ensemble = False

    @property
    def runtime(self):
        rt = super().runtime
        rt.want_local = True
        return rt

    def _dirs(self):
        return ("f", "b") if self.bidirectional else ("f",)

    @torch.no_grad()
    def dense_local(self, res, direction: str = "f"):
        """``hist_embeddings_loc [B, M, D]`` (PostDynamicRGCN.py:33-42): the local stream of the last history step that ran,
        fresh zeros elsewhere ("history forgets")."""
        loc = self.ent_embeds.new_zeros(res.plan.batch, self.num_ents, self.embed_size)
        lasts = res.plan.last_hist_f if direction == "f" else res.plan.last_hist_b
        for i, inst in enumerate(lasts):
            if inst is not None:
                ids = torch.from_numpy(inst.snapshot.node_ids).to(loc.device)
                loc[i][ids] = res.bufs["local"][inst.row0:inst.row0 + inst.n]
        return loc

    def _window(self, t_list):
        """encode + the reference-shaped dense tensors per direction: [(loc, rec, start), ...]."""
        res = self.encode(t_list)
        self.last_result = res
        dense = []
        for d in self._dirs():
            rec, start = self.dense_history(res, d)
            dense.append((self.dense_local(res, d), rec, start))
        fin = res.plan.final
        local = res.bufs["local"][fin.row0:fin.row1].clone()
        return res, dense, tuple(local.split(res.plan.final_sizes)), tuple(res.out.clone().split(res.plan.final_sizes))

    def _iso_args(self, dense, i, cur_t):
        """The dense argument list of the encoder's isolated calls for batch item i (both directions for Bi)."""
        args, locs = [], []
        for loc, rec, start in dense:
            args += [rec[i][0], rec[i][1], (cur_t - start[i]).unsqueeze(-1)]
            locs.append(loc[i])
        return args, locs

    def _graph_rows(self, table, g, rows):
        table[torch.from_numpy(g.node_ids).to(table.device)] = rows
        return table

    def forward(self, t_list, reverse=False):
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise NotImplementedError("temp_b200: training the post-ensemble / impute variants is not built (forward-only "
                                      "shells); wrap the call in torch.no_grad() for the loss value")
        with torch.no_grad():
            return self._loss(t_list)

    def evaluate(self, t_list, val=True):
        with torch.no_grad():
            return self._evaluate_from(self.evaluate_embed(t_list, val))


class ImputeDynamicRGCN(_PostBase, DynamicRGCN):
    """models/PostDynamicRGCN.py:20-143."""

    @torch.no_grad()
    def get_all_embeds_Gt(self, convoluted_embeds, g, t, second_embeds_loc, first_prev_graph_embeds, second_prev_graph_embeds,
                          time_diff_tensor):
        """PostDynamicRGCN.py:24-31."""
        alls = self.ent_encoder.forward_isolated_impute(self.ent_embeds, first_prev_graph_embeds, second_prev_graph_embeds,
                                                        time_diff_tensor.unsqueeze(-1), t, second_embeds_loc)
        return self._graph_rows(alls, g, convoluted_embeds)

    @torch.no_grad()
    def evaluate_embed(self, t_list, val=True):
        """PostDynamicRGCN.py:101-113 -> (per_graph, test_graphs, time_list, hist_loc, hist_rec, start_time_tensor)."""
        res, dense, _, rec = self._window(t_list)
        graph_dict = self.graph_dict_val if val else self.graph_dict_test
        _, time_list = self.get_batch_graph_list(t_list, self.test_seq_len, self.graph_dict_train)
        loc_h, rec_h, start = dense[0]
        return rec, [graph_dict.get(t) for t in res.plan.final_times], time_list, loc_h, rec_h, start

    def _evaluate_from(self, out):
        rec, graphs, time_list, loc_h, rec_h, start = out
        return self.calc_metrics(rec, graphs, time_list[-1], loc_h, rec_h, start, self.test_seq_len - 1)

    @torch.no_grad()
    def calc_metrics(self, per_graph_ent_embeds, g_list, t_list, hist_embeddings_loc, hist_embeddings_rec, start_time_tensor, cur_t):
        """PostDynamicRGCN.py:119-143 (the history index only advances for graphs with edges)."""
        self._ensure_evaluater()
        dev = self.ent_embeds.device
        ranks, losses = [], []
        i = 0
        for g, t, ent_embed in zip(g_list, t_list, per_graph_ent_embeds):
            alls = self.get_all_embeds_Gt(ent_embed, g, t, hist_embeddings_loc[i], hist_embeddings_rec[i][0],
                                          hist_embeddings_rec[i][1], cur_t - start_time_tensor[i])
            if g is None or g.num_edges == 0:
                continue
            src, dst = g.edges()
            index_sample = torch.stack([src, g.edata["type_s"], dst]).transpose(0, 1).to(dev)
            ranks.append(self.evaluater.calc_metrics_single_graph(ent_embed, self.rel_embeds, alls, index_sample, g, t))
            losses.append(self.link_classification_loss(ent_embed, self.rel_embeds, index_sample,
                                                        torch.ones(index_sample.shape[0], device=dev)).item())
            i += 1
        ranks = torch.cat(ranks) if ranks else torch.zeros(0, dtype=torch.long, device=dev)
        return ranks, (float(np.mean(losses)) if losses else float("nan"))

    def _ensure_evaluater(self):
        from .evaluation import EvaluationFilter
        if getattr(self, "evaluater", None) is None:
            self.evaluater = EvaluationFilter(self.args, self.calc_score, self.graph_dict_train, self.graph_dict_val,
                                              self.graph_dict_test)

    def _loss(self, t_list):
        """PostDynamicRGCN.py:81-99 (the window runs on full graphs in eval mode, edge-sub-sampled in train mode)."""
        if self.training:
            raise NotImplementedError("temp_b200: the train-mode (edge sub-sampled) loss of the impute variants is not built")
        res, dense, _, rec = self._window(t_list)
        dev = self.ent_embeds.device
        L = self.train_seq_len
        loss = 0
        for i, (t, g, ent_embed) in enumerate(zip(res.plan.final_times, res.plan.final_snapshots, rec)):
            tri, neg_t, neg_h, labels = (x.to(dev) for x in self.corrupter.single_graph_negative_sampling(t, g, self.num_ents))
            args, locs = self._iso_args(dense, i, L - 1)
            alls = self._all_impute(ent_embed, g, t, args, locs)
            loss = loss + self.train_link_prediction(ent_embed, tri, neg_t, labels, alls, corrupt_tail=True)
            loss = loss + self.train_link_prediction(ent_embed, tri, neg_h, labels, alls, corrupt_tail=False)
        return loss

    def _all_impute(self, ent_embed, g, t, args, locs):
        return self.get_all_embeds_Gt(ent_embed, g, t, locs[0], args[0], args[1], args[2].squeeze(-1))


class PostEnsembleDynamicRGCN(ImputeDynamicRGCN):
    """models/PostDynamicRGCN.py:324-462 (on top of PostDynamicRGCN 146-322)."""
    post_ensemble = True

    def build_model(self):
        super().build_model()
        self.subject_linear, self.object_linear = _freq_mlp(), _freq_mlp()
        self.drop_edge = FrequencyStats(self.graph_dict_train, self.train_seq_len, self.bidirectional)

    @torch.no_grad()
    def get_all_embeds_Gt(self, convoluted_embeds_loc, convoluted_embeds, g, t, second_embeds_loc, first_prev_graph_embeds,
                          second_prev_graph_embeds, time_diff_tensor):
        """PostDynamicRGCN.py:174-187 -> (all_embeds_g_loc, all_embeds_g_rec)."""
        loc, rec = self.ent_encoder.forward_post_ensemble_isolated(self.ent_embeds, first_prev_graph_embeds,
                                                                   second_prev_graph_embeds, time_diff_tensor.unsqueeze(-1), t,
                                                                   second_embeds_loc)
        return self._graph_rows(loc.clone(), g, convoluted_embeds_loc), self._graph_rows(rec, g, convoluted_embeds)

    @torch.no_grad()
    def evaluate_embed(self, t_list, val=True):
        """PostDynamicRGCN.py:371-383 -> (per_graph_loc, per_graph_rec, test_graphs, time_list, hist_loc, hist_rec, start)."""
        res, dense, loc, rec = self._window(t_list)
        graph_dict = self.graph_dict_val if val else self.graph_dict_test
        _, time_list = self.get_batch_graph_list(t_list, self.test_seq_len, self.graph_dict_train)
        loc_h, rec_h, start = dense[0]
        return loc, rec, [graph_dict.get(t) for t in res.plan.final_times], time_list, loc_h, rec_h, start

    def _evaluate_from(self, out):
        loc, rec, graphs, time_list, loc_h, rec_h, start = out
        return self.calc_metrics(loc, rec, graphs, time_list[-1], loc_h, rec_h, start, self.test_seq_len - 1)

    @torch.no_grad()
    def calc_ensemble_ratio(self, triples, t, g):
        """PostDynamicRGCN.py:430-462 -> (weight_subject, weight_object), each [n, 1]."""
        dev = self.ent_embeds.device
        ids = g.node_ids
        tri = [(int(ids[s]), int(r), int(ids[o])) for s, r, o in triples.tolist()]
        sub_f, obj_f = self.drop_edge.features(tri, int(t))
        return torch.sigmoid(self.subject_linear(sub_f.to(dev))), torch.sigmoid(self.object_linear(obj_f.to(dev)))

    def _all_post(self, loc_i, rec_i, g, t, args, locs):
        return self.get_all_embeds_Gt(loc_i, rec_i, g, t, locs[0], args[0], args[1], args[2].squeeze(-1))

    @torch.no_grad()
    def calc_metrics(self, per_graph_ent_embeds_loc, per_graph_ent_embeds_rec, g_list, t_list, hist_embeddings_loc,
                     hist_embeddings_rec, start_time_tensor, cur_t):
        """PostDynamicRGCN.py:385-411."""
        dense = [(hist_embeddings_loc, hist_embeddings_rec, start_time_tensor)]
        return self._ensemble_metrics(per_graph_ent_embeds_loc, per_graph_ent_embeds_rec, g_list, t_list, dense, cur_t)

    def _ensemble_metrics(self, locs_pg, recs_pg, g_list, t_list, dense, cur_t):
        self._ensure_evaluater()
        dev = self.ent_embeds.device
        ranks = []
        i = 0
        for g, t, loc_i, rec_i in zip(g_list, t_list, locs_pg, recs_pg):
            args, locs = self._iso_args(dense, i, cur_t)
            all_loc, all_rec = self._all_post(loc_i, rec_i, g, t, args, locs)
            if g is None or g.num_edges == 0:
                continue
            src, dst = g.edges()
            samples = torch.stack([src, g.edata["type_s"], dst]).transpose(0, 1).to(dev)
            w_sub, w_obj = self.calc_ensemble_ratio(samples, t, g)
            ranks.append(ensemble_ranks(self.evaluater, self.calc_score, loc_i, rec_i, self.rel_embeds, all_loc, all_rec,
                                        w_sub, w_obj, samples, g, t))
            i += 1
        ranks = torch.cat(ranks) if ranks else torch.zeros(0, dtype=torch.long, device=dev)
        return ranks, float("nan")                      # (the reference returns np.mean([]) here: no loss is computed)

    def _loss(self, t_list):
        """PostDynamicRGCN.py:340-365: cross-entropy of the weighted sum of the local and the recurrent scores."""
        if self.training:
            raise NotImplementedError("temp_b200: the train-mode (edge sub-sampled) loss of the post-ensemble variants is not built")
        res, dense, loc, rec = self._window(t_list)
        dev = self.ent_embeds.device
        L = self.train_seq_len
        loss = 0
        for i, (t, g) in enumerate(zip(res.plan.final_times, res.plan.final_snapshots)):
            tri, neg_t, neg_h, labels = (x.to(dev) for x in self.corrupter.single_graph_negative_sampling(t, g, self.num_ents))
            args, locs = self._iso_args(dense, i, L - 1)
            all_loc, all_rec = self._all_post(loc[i], rec[i], g, t, args, locs)
            w_sub, w_obj = self.calc_ensemble_ratio(tri, t, g)
            r = self.rel_embeds[tri[:, 1]]

            def scores(ent, alls, neg, tail):
                if tail:
                    return self.calc_score(ent[tri[:, 0]], r, alls[neg], mode="tail")
                return self.calc_score(alls[neg], r, ent[tri[:, 2]], mode="head")
            tail = w_obj * scores(loc[i], all_loc, neg_t, True) + (1 - w_obj) * scores(rec[i], all_rec, neg_t, True)
            head = w_sub * scores(loc[i], all_loc, neg_h, False) + (1 - w_sub) * scores(rec[i], all_rec, neg_h, False)
            loss = loss + F.cross_entropy(tail, labels) + F.cross_entropy(head, labels)
        return loss


class ImputeBiDynamicRGCN(_PostBase, BiDynamicRGCN):
    """models/PostBiDynamicRGCN.py:23-175."""

    @torch.no_grad()
    def get_all_embeds_Gt(self, convoluted_embeds, g, t, second_embeds_forward_loc, first_prev_graph_embeds_forward_rec,
                          second_prev_graph_embeds_forward_rec, time_diff_tensor_forward, second_embeds_backward_loc,
                          first_prev_graph_embeds_backward_rec, second_prev_graph_embeds_backward_rec, time_diff_tensor_backward):
        """PostBiDynamicRGCN.py:30-40."""
        alls = self.ent_encoder.forward_isolated_impute(
            self.ent_embeds, first_prev_graph_embeds_forward_rec, second_prev_graph_embeds_forward_rec,
            time_diff_tensor_forward.unsqueeze(-1), first_prev_graph_embeds_backward_rec, second_prev_graph_embeds_backward_rec,
            time_diff_tensor_backward.unsqueeze(-1), t, second_embeds_forward_loc, second_embeds_backward_loc)
        return self._graph_rows(alls, g, convoluted_embeds)

    @torch.no_grad()
    def evaluate_embed(self, t_list, val=True):
        """PostBiDynamicRGCN.py:124-142 -> (per_graph, test_graphs, times, loc_f, rec_f, start_f, loc_b, rec_b, start_b)."""
        res, dense, _, rec = self._window(t_list)
        graph_dict = self.graph_dict_val if val else self.graph_dict_test
        return (rec, [graph_dict.get(t) for t in res.plan.final_times], list(res.plan.final_times)) + dense[0] + dense[1]

    def _evaluate_from(self, out):
        return self.calc_metrics(*out, self.test_seq_len - 1)

    _ensure_evaluater = ImputeDynamicRGCN._ensure_evaluater

    @torch.no_grad()
    def calc_metrics(self, per_graph_ent_embeds, g_list, t_list, hist_embeddings_forward_loc, hist_embeddings_forward_rec,
                     start_time_tensor_forward, hist_embeddings_backward_loc, hist_embeddings_backward_rec,
                     start_time_tensor_backward, cur_t):
        """PostBiDynamicRGCN.py:150-175."""
        self._ensure_evaluater()
        dev = self.ent_embeds.device
        dense = [(hist_embeddings_forward_loc, hist_embeddings_forward_rec, start_time_tensor_forward),
                 (hist_embeddings_backward_loc, hist_embeddings_backward_rec, start_time_tensor_backward)]
        ranks, losses = [], []
        i = 0
        for g, t, ent_embed in zip(g_list, t_list, per_graph_ent_embeds):
            args, locs = self._iso_args(dense, i, cur_t)
            alls = self._all_impute(ent_embed, g, t, args, locs)
            if g is None or g.num_edges == 0:
                continue
            src, dst = g.edges()
            index_sample = torch.stack([src, g.edata["type_s"], dst]).transpose(0, 1).to(dev)
            ranks.append(self.evaluater.calc_metrics_single_graph(ent_embed, self.rel_embeds, alls, index_sample, g, t))
            losses.append(self.link_classification_loss(ent_embed, self.rel_embeds, index_sample,
                                                        torch.ones(index_sample.shape[0], device=dev)).item())
            i += 1
        ranks = torch.cat(ranks) if ranks else torch.zeros(0, dtype=torch.long, device=dev)
        return ranks, (float(np.mean(losses)) if losses else float("nan"))

    def _all_impute(self, ent_embed, g, t, args, locs):
        return self.get_all_embeds_Gt(ent_embed, g, t, locs[0], args[0], args[1], args[2].squeeze(-1),
                                      locs[1], args[3], args[4], args[5].squeeze(-1))

    _loss = ImputeDynamicRGCN._loss


class PostEnsembleBiDynamicRGCN(ImputeBiDynamicRGCN):
    """models/PostBiDynamicRGCN.py:178-372."""
    post_ensemble = True

    def build_model(self):
        super().build_model()
        self.subject_linear, self.object_linear = _freq_mlp(), _freq_mlp()
        self.drop_edge = FrequencyStats(self.graph_dict_train, self.train_seq_len, True)

    @torch.no_grad()
    def get_all_embeds_Gt(self, convoluted_embeds_loc, convoluted_embeds_rec, g, t, second_embeds_forward_loc,
                          first_prev_graph_embeds_forward_rec, second_prev_graph_embeds_forward_rec, time_diff_tensor_forward,
                          second_embeds_backward_loc, first_prev_graph_embeds_backward_rec, second_prev_graph_embeds_backward_rec,
                          time_diff_tensor_backward):
        """PostBiDynamicRGCN.py:184-195 -> (all_embeds_g_loc, all_embeds_g_rec)."""
        loc, rec = self.ent_encoder.forward_post_ensemble_isolated(
            self.ent_embeds, first_prev_graph_embeds_forward_rec, second_prev_graph_embeds_forward_rec,
            time_diff_tensor_forward.unsqueeze(-1), first_prev_graph_embeds_backward_rec, second_prev_graph_embeds_backward_rec,
            time_diff_tensor_backward.unsqueeze(-1), t, second_embeds_forward_loc, second_embeds_backward_loc)
        return self._graph_rows(loc.clone(), g, convoluted_embeds_loc), self._graph_rows(rec, g, convoluted_embeds_rec)

    @torch.no_grad()
    def evaluate_embed(self, t_list, val=True):
        """PostBiDynamicRGCN.py:227-245 -> (per_graph_loc, per_graph_rec, test_graphs, times, loc_f, rec_f, start_f, loc_b,
        rec_b, start_b)."""
        res, dense, loc, rec = self._window(t_list)
        graph_dict = self.graph_dict_val if val else self.graph_dict_test
        return (loc, rec, [graph_dict.get(t) for t in res.plan.final_times], list(res.plan.final_times)) + dense[0] + dense[1]

    calc_ensemble_ratio = PostEnsembleDynamicRGCN.calc_ensemble_ratio
    _ensemble_metrics = PostEnsembleDynamicRGCN._ensemble_metrics
    _loss = PostEnsembleDynamicRGCN._loss

    def _all_post(self, loc_i, rec_i, g, t, args, locs):
        return self.get_all_embeds_Gt(loc_i, rec_i, g, t, locs[0], args[0], args[1], args[2].squeeze(-1),
                                      locs[1], args[3], args[4], args[5].squeeze(-1))

    @torch.no_grad()
    def calc_metrics(self, per_graph_ent_embeds_loc, per_graph_ent_embeds_rec, g_list, t_list, hist_embeddings_forward_loc,
                     hist_embeddings_forward_rec, start_time_tensor_forward, hist_embeddings_backward_loc,
                     hist_embeddings_backward_rec, start_time_tensor_backward, cur_t):
        """PostBiDynamicRGCN.py:296-320."""
        dense = [(hist_embeddings_forward_loc, hist_embeddings_forward_rec, start_time_tensor_forward),
                 (hist_embeddings_backward_loc, hist_embeddings_backward_rec, start_time_tensor_backward)]
        return self._ensemble_metrics(per_graph_ent_embeds_loc, per_graph_ent_embeds_rec, g_list, t_list, dense, cur_t)


@torch.no_grad()
def ensemble_ranks(evaluater, calc_score, ent_loc, ent_rec, rel, all_loc, all_rec, w_sub, w_obj, samples, graph, t):
    """utils/post_evaluation.py:82-134 (PostEnsembleEvaluationFilter): per side the masked local and recurrent scores against
    all entities, ``w * local + (1 - w) * recurrent`` -- the OBJECT side (mode 'tail') takes weight_subject, the subject side
    weight_object (lines 91-94) --, sigmoid, stable descending sort, position of the target; subject side first, 1-indexed."""
    dev = all_loc.device
    samples_cpu = samples.cpu()
    ids = torch.from_numpy(graph.node_ids).to(dev)
    r = rel[samples[:, 1]]
    out = {}
    for mode, w in (("tail", w_sub), ("head", w_obj)):
        mask = evaluater._mask(samples_cpu, all_loc.shape[0], t, graph, mode).to(dev)
        scs = []
        for ent, alls in ((ent_loc, all_loc), (ent_rec, all_rec)):
            if mode == "tail":
                sc = calc_score(ent[samples[:, 0]], r, alls, mode="tail")
            else:
                sc = calc_score(alls, r, ent[samples[:, 2]], mode="head")
            scs.append(torch.where(mask, torch.full_like(sc, -10e6), sc))
        target = ids[samples[:, 2 if mode == "tail" else 0]]
        sc = torch.sigmoid(w * scs[0] + (1 - w) * scs[1])
        _, order = torch.sort(sc, dim=1, descending=True, stable=True)
        out[mode] = torch.nonzero(order == target.view(-1, 1))[:, 1].view(-1)
    return torch.cat([out["head"], out["tail"]]) + 1
