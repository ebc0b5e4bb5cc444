"""Filtered ranking of the reference (utils/evaluation.py:6-106) -- a "next" row
kept in torch: scores of every query against all entities in batches of 100, known true answers of
the same (timestamp, query) masked out, rank of the target by a descending sort."""
from __future__ import annotations

import numpy as np
import torch

from .sampler import CorruptTriples


class EvaluationFilter(object):
    def __init__(self, args, calc_score, graph_dict_train, graph_dict_val, graph_dict_test):
        self.args, self.calc_score = args, calc_score
        self.graph_dict_train, self.graph_dict_val, self.graph_dict_test = graph_dict_train, graph_dict_val, graph_dict_test
        self.true_heads, self.true_tails = {}, {}

    def _true_sets(self, t):
        if t not in self.true_heads:
            parts = []
            for gd in (self.graph_dict_train, self.graph_dict_val, self.graph_dict_test):
                g = gd.get(t)
                if g is not None:
                    parts.append(np.stack([g.src, g.rel, g.dst], axis=1))
            th, tt = CorruptTriples.get_true_head_and_tail_per_graph(np.concatenate(parts, axis=0))
            self.true_heads[t], self.true_tails[t] = th, tt
        return self.true_heads[t], self.true_tails[t]

    def _mask(self, samples, num_ent, t, graph, mode):
        th, tt = self._true_sets(int(t))
        ids = graph.node_ids
        mask = torch.zeros(samples.shape[0], num_ent, dtype=torch.bool)
        for i, (h, r, tl) in enumerate(samples.tolist()):
            if mode == "tail":
                mask[i, torch.from_numpy(ids[tt[(h, r)]])] = True
                mask[i, int(ids[tl])] = False
            else:
                mask[i, torch.from_numpy(ids[th[(r, tl)]])] = True
                mask[i, int(ids[h])] = False
        return mask

    def calc_metrics_single_graph(self, ent_mean, rel_enc_means, all_ent_embeds, samples, graph, time, eval_bz=100):
        with torch.no_grad():
            dev = all_ent_embeds.device
            samples_cpu = samples.cpu()
            num_ent = all_ent_embeds.shape[0]
            ids = torch.from_numpy(graph.node_ids).to(dev)
            out = []
            for mode in ("head", "tail"):                             # ranks_s first, then ranks_o (line 48)
                mask = self._mask(samples_cpu, num_ent, time, graph, mode).to(dev)
                ranks = []
                for lo in range(0, samples.shape[0], eval_bz):
                    sl = slice(lo, min(samples.shape[0], lo + eval_bz))
                    r = rel_enc_means[samples[sl, 1]]
                    if mode == "tail":
                        score = self.calc_score(ent_mean[samples[sl, 0]], r, all_ent_embeds, mode=mode)
                        target = ids[samples[sl, 2]]
                    else:
                        score = self.calc_score(all_ent_embeds, r, ent_mean[samples[sl, 2]], mode=mode)
                        target = ids[samples[sl, 0]]
                    score = torch.where(mask[sl], torch.full_like(score, -10e6), score)
                    score = torch.sigmoid(score)
                    _, order = torch.sort(score, dim=1, descending=True)
                    ranks.append(torch.nonzero(order == target.view(-1, 1))[:, 1].view(-1))
                out.append(torch.cat(ranks))
            return torch.cat(out) + 1
