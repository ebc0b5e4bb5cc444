"""Filtered negative sampling, host-side and BIT-EXACT with the reference (utils/CorrptTriples.py).

Not accelerated on purpose: the indices depend on the global NumPy / torch
RNG streams, so the call sequence is preserved exactly --
  per graph : ``torch.randperm(E)`` iff E > num_pos_facts                       (CorrptTriples.py:36-40)
  per triple: tail corruption rounds, then head corruption rounds, each round one
              ``np.random.randint(num_entities, size=negative_rate)`` filtered by
              ``np.in1d(cand, true_global_ids, assume_unique=True, invert=True)``  (CorrptTriples.py:61-85)
Column 0 of both sample matrices is the global id of the true entity; labels are all zero.
"""
from __future__ import annotations

import numpy as np
import torch


class CorruptTriples(object):
    def __init__(self, args, graph_dict_train):
        self.args = args
        self.negative_rate = args.negative_rate
        self.num_pos_facts = args.num_pos_facts
        self.graph_dict_train = graph_dict_train
        self._true = {}

    @staticmethod
    def get_true_head_and_tail_per_graph(triples: np.ndarray):
        """(relation, tail) -> heads and (head, relation) -> tails, local node ids (CorrptTriples.py:87-106)."""
        heads, tails = {}, {}
        for h, r, t in triples.tolist():
            tails.setdefault((h, r), []).append(t)
            heads.setdefault((r, t), []).append(h)
        heads = {k: np.array(list(set(v))) for k, v in heads.items()}
        tails = {k: np.array(list(set(v))) for k, v in tails.items()}
        return heads, tails

    def _true_sets(self, t, g):
        hit = self._true.get(t)
        if hit is None:
            hit = self.get_true_head_and_tail_per_graph(np.stack([g.src, g.rel, g.dst], axis=1))
            self._true[t] = hit
        return hit

    def _corrupt(self, forbidden_global, num_entities):
        chunks, size = [], 0
        while size < self.negative_rate:
            cand = np.random.randint(num_entities, size=self.negative_rate)
            cand = cand[np.isin(cand, forbidden_global, assume_unique=True, invert=True)]
            chunks.append(cand)
            size += cand.size
        return np.concatenate(chunks)[:self.negative_rate]

    def single_graph_negative_sampling(self, t, g, num_ents):
        """-> (triples LongTensor [P,3] local ids, neg_tail [P,1+neg], neg_head [P,1+neg], labels [P])."""
        t = int(t)
        true_head, true_tail = self._true_sets(t, g)
        triples = np.stack([g.src, g.rel, g.dst], axis=1)
        P = min(triples.shape[0], self.num_pos_facts)
        if self.num_pos_facts < triples.shape[0]:
            perm = torch.randperm(triples.shape[0]).numpy()
            triples = triples[perm[:self.num_pos_facts]]
        neg_tail = np.zeros((P, 1 + self.negative_rate), dtype=int)
        neg_head = np.zeros((P, 1 + self.negative_rate), dtype=int)
        node_ids = g.node_ids
        for i in range(P):
            h, r, tl = (int(x) for x in triples[i])
            tail_s = self._corrupt([int(node_ids[j]) for j in true_tail[(h, r)].tolist()], num_ents)
            head_s = self._corrupt([int(node_ids[j]) for j in true_head[(r, tl)].tolist()], num_ents)
            neg_tail[i, 0], neg_head[i, 0] = node_ids[tl], node_ids[h]
            neg_tail[i, 1:], neg_head[i, 1:] = tail_s, head_s
        labels = np.zeros(P, dtype=int)
        return (torch.from_numpy(triples), torch.from_numpy(neg_tail), torch.from_numpy(neg_head),
                torch.from_numpy(labels))
