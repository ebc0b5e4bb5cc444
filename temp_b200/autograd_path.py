"""Differentiable statement of the DynamicRGCN training forward, used ONLY when gradients are requested.

SURVEY.md section 8b ("Autograd"): the reference trains through this path (``loss.backward()`` from Lightning); the CUDA
kernels of this round are forward-only, so ``forward()`` with gradients enabled falls back to the torch operators below
(on the GPU, through torch autograd -- NOT accelerated, documented as such) so that ``main.py``-style training runs:

    model.train(); loss = model(t_list); loss.backward(); optimizer.step()

It restates, on the packed window plan of temp_b200/planner.py (same ``prev_row`` history maps as the kernels):
  RGCNLayer.forward / forward_isolated        models/RGCN.py:53-104 (self-loop dropout included, RGCN.py:58-59, 84)
  GRRGCNLayer / RRGCNLayer (rec-only layer 2)  models/RRGCN.py:77-89, 130-154; forward_isolated 91-104, 156-167
  RRGCN.forward / forward_isolated            models/RRGCN.py:192-217
  DynamicRGCN.forward, get_all_embeds_Gt      models/DynamicRGCN.py:56-64, 176-194
Scope: GRRGCN / RRGCN with ``--rec-only-last-layer`` (the shipped uni-directional configurations); everything else
raises.  The no-grad forward never comes here (it is the CUDA path and has no fallback).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def _check(model):
    if model.ent_embeds.device.type != "cuda":
        raise RuntimeError("temp_b200: the model must live on a CUDA device")
    if model.family != "recurrent" or model.bidirectional or not model.ent_encoder.rec_only_last_layer:
        raise NotImplementedError("temp_b200: the autograd fallback covers GRRGCN / RRGCN with --rec-only-last-layer; "
                                  "use torch.no_grad() for the forward of this configuration")


class _Ctx(object):
    """Device copies of the plan arrays the torch statement needs."""

    def __init__(self, model, plan):
        dev = model.ent_embeds.device
        t = lambda a, dt=torch.long: torch.as_tensor(np.ascontiguousarray(a), device=dev).to(dt)
        self.plan = plan
        self.ent_id, self.row_time = t(plan.ent_id), t(plan.row_time)
        self.norm = t(plan.norm, torch.float32)
        deg = np.diff(plan.row_ptr.astype(np.int64))
        self.e_dst = t(np.repeat(np.arange(plan.R, dtype=np.int64), deg))
        self.e_src, self.e_rel = t(plan.e_src), t(plan.e_rel)
        self.prev, self.dt = t(plan.prev_a), t(plan.dt_a, torch.float32)


def _rgcn_graph(layer, h, c: _Ctx, training: bool):
    """models/RGCN.py:53-104: block-diagonal messages, edge norm, sum, node norm, (+bias), + dropout(self loop), act."""
    E, D = c.e_src.shape[0], h.shape[1]
    nb, si, so = layer.num_bases, layer.submat_in, layer.submat_out
    w = layer.weight.index_select(0, c.e_rel).view(-1, si, so)
    msg = torch.bmm(h.index_select(0, c.e_src).view(-1, 1, si), w).view(E, D)
    msg = msg * c.norm.index_select(0, c.e_dst).unsqueeze(1)
    agg = torch.zeros(h.shape[0], D, dtype=h.dtype, device=h.device).index_add(0, c.e_dst, msg) * c.norm.unsqueeze(1)
    loop = F.dropout(h @ layer.loop_weight, p=layer.dropout_p, training=training)
    out = agg + (layer.h_bias if layer.bias else 0) + loop
    return torch.relu(out) if layer.activation == "relu" else out


def _rgcn_isolated(layer, x, training: bool):
    """models/RGCN.py:78-89 (note the residual)."""
    out = x + F.dropout(x @ layer.loop_weight, p=layer.dropout_p, training=training)
    if layer.bias:
        out = out + layer.h_bias
    return torch.relu(out) if layer.activation == "relu" else out


def _decay(model, layer, dt):
    if layer.learnable_lambda:                                   # models/RGCN.py:106-107
        return torch.exp(-torch.clamp(layer.exponential_decay(dt.view(-1, 1)), min=0))
    return torch.exp(-dt.view(-1, 1) * float(model.args.inv_temperature))


def _gru(model, rnn, x, h0):
    """torch.nn.GRU with sequence length 1 (models/RRGCN.py:84) or the --type1 cell (models/GRU_cell.py:18-31)."""
    if bool(getattr(model.args, "type1", False)):
        D = h0.shape[1]
        i_n = x @ rnn.weight_ih.t() + rnn.bias_ih
        gh = h0 @ rnn.weight_hh.t() + rnn.bias_hh
        r, z = torch.sigmoid(gh[:, :D]), torch.sigmoid(gh[:, D:2 * D])
        n = torch.tanh(i_n + r * gh[:, 2 * D:])
        return n + z * (h0 - n)
    _, hn = rnn(x.unsqueeze(0), h0.unsqueeze(0).contiguous())
    return hn.squeeze(0)


def encode(model, plan):
    """-> (state [R, D] of every packed row, ctx); differentiable w.r.t. the model parameters."""
    _check(model)
    c = _Ctx(model, plan)
    enc = model.ent_encoder
    l1, l2 = enc.layer_1, enc.layer_2
    training = model.training
    gru = model.args.module == "GRRGCN"
    h1 = _rgcn_graph(l1, model.ent_embeds.index_select(0, c.ent_id), c, training)
    S = torch.zeros(plan.R, model.embed_size, dtype=h1.dtype, device=h1.device)
    x2 = _rgcn_graph(l2, h1, c, training)       # layer 2 of the uni-directional models has no activation (RRGCN.py:186-187)
    for seg in plan.segments:
        rows = torch.arange(seg.row0, seg.row1, device=S.device)
        prev = c.prev.index_select(0, rows)
        has = (prev >= 0).unsqueeze(1)
        raw = torch.where(has, S.index_select(0, prev.clamp(min=0)), torch.zeros((), device=S.device))
        dec = _decay(model, l2, c.dt.index_select(0, rows))
        x = x2.index_select(0, rows)
        if gru:
            hn = _gru(model, l2.rnn, x, raw * dec)                                     # RRGCN.py:79-85
        else:
            hn = x + (raw @ l2.time_weight) * torch.exp(-c.dt.index_select(0, rows).view(-1, 1)
                                                        * float(model.args.inv_temperature))  # RRGCN.py:142
        if enc.use_time_embedding:
            hn = hn + l2.time_embed.index_select(0, c.row_time.index_select(0, rows))   # RRGCN.py:202-203
        S = S.index_copy(0, rows, hn)
    return S, c


def all_embeds(model, plan, S, i: int):
    """get_all_embeds_Gt (models/DynamicRGCN.py:56-64): forward_isolated over all entities with item i's history, rows
    of the target graph's entities overwritten by the graph states."""
    enc = model.ent_encoder
    l1, l2 = enc.layer_1, enc.layer_2
    training = model.training
    M, D, L = model.num_ents, model.embed_size, plan.seq_len
    dev = S.device
    hist = torch.zeros(M, D, dtype=S.dtype, device=dev)
    start = torch.zeros(M, dtype=S.dtype, device=dev)
    for seg in plan.segments:                                    # start_time: last history step an entity was active in
        if seg.kind != "hist_f":
            continue
        for inst in seg.instances:
            if inst.item == i:
                start[torch.as_tensor(inst.snapshot.node_ids, device=dev)] = float(inst.step)
    last = plan.last_hist_f[i]
    if last is not None:                                         # "history forgets": only the last step's entities
        ids = torch.as_tensor(last.snapshot.node_ids, device=dev)
        hist = hist.index_copy(0, ids, S[last.row0:last.row0 + last.n])
    dt = (L - 1) - start
    t = plan.final_times[i]
    first = _rgcn_isolated(l1, model.ent_embeds, training)
    if model.args.module == "GRRGCN":
        x = _rgcn_isolated(l2, first, training)
        second = _gru(model, l2.rnn, x, hist * _decay(model, l2, dt))                   # RRGCN.py:91-98
    else:
        second = first + F.dropout(first @ l2.loop_weight, p=l2.dropout_p, training=training)
        second = second + (hist @ l2.time_weight) * torch.exp(-dt.view(-1, 1) * float(model.args.inv_temperature))
    if enc.use_time_embedding:
        second = second + l2.time_embed[int(t)]
    fin = plan.final.instances[i]
    ids = torch.as_tensor(fin.snapshot.node_ids, device=dev)
    return second.index_copy(0, ids, S[fin.row0:fin.row0 + fin.n])


def training_loss(model, t_list):
    """models/DynamicRGCN.py:176-194 with autograd: sub-sampled window in train() mode, bit-exact negative sampling,
    tail + head cross-entropy through the torch scorers."""
    _check(model)
    plan = model.plan(t_list, transform=model.train_edge_sampler() if model.training else None)
    S, _ = encode(model, plan)
    dev = S.device
    loss = 0
    for i, (t, g) in enumerate(zip(plan.final_times, plan.final_snapshots)):
        fin = plan.final.instances[i]
        ent_embed = S[fin.row0:fin.row0 + fin.n]
        triplets, neg_tail, neg_head, labels = model.corrupter.single_graph_negative_sampling(t, g, model.num_ents)
        triplets, neg_tail, neg_head, labels = (x.to(dev) for x in (triplets, neg_tail, neg_head, labels))
        all_g = all_embeds(model, plan, S, i)
        loss = loss + model.train_link_prediction(ent_embed, triplets, neg_tail, labels, all_g, corrupt_tail=True)
        loss = loss + model.train_link_prediction(ent_embed, triplets, neg_head, labels, all_g, corrupt_tail=False)
    return loss
