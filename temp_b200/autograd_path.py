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
  BiGRRGCNLayer / BiRRGCNLayer, BiRRGCN       models/BiRRGCN.py:27-81, 115-186, 188-257
  BiDynamicRGCN.forward, get_all_embeds_Gt    models/BiDynamicRGCN.py:102-112, 165-187
  RGCN (static)                               models/RGCN.py:154-164, baselines/StaticRGCN.py:36-89
  SARGCNLayer / SARGCN, (Bi)SelfAttentionRGCN models/SARGCN.py:25-62, 103-125; models/SelfAttentionRGCN.py:26-43, 73-139;
                                              models/BiSelfAttentionRGCN.py:17-87
Scope: every family of the module table (GRRGCN / RRGCN / BiGRRGCN / BiRRGCN with or without ``--rec-only-last-layer``, torch
GRU or --type1 cell, learnable lambda; SRGCN; SARGCN / BiSARGCN); the post-ensemble / impute shells raise.  The no-grad
forward never comes here (it is the CUDA path and has no fallback).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def _check(model):
    if model.ent_embeds.device.type != "cuda":
        raise RuntimeError("temp_b200: the model must live on a CUDA device")


class _Ctx(object):
    """Device copies of the plan arrays the torch statement needs."""

    def __init__(self, model, plan):
        dev = model.ent_embeds.device
        t = lambda a, dt=torch.long: torch.as_tensor(np.ascontiguousarray(a), device=dev).to(dt)
        self.plan = plan
        self.ent_id, self.row_time = t(plan.ent_id), t(plan.row_time)
        self.norm = t(plan.norm, torch.float32)
        self.row_ptr = plan.row_ptr.astype(np.int64)
        deg = np.diff(self.row_ptr)
        self.e_dst = t(np.repeat(np.arange(plan.R, dtype=np.int64), deg))
        self.e_src, self.e_rel = t(plan.e_src), t(plan.e_rel)
        self.prev = {"a": t(plan.prev_a)}
        self.dt = {"a": t(plan.dt_a, torch.float32)}
        if plan.bidirectional:                                   # second chain of the centre step (backward history)
            self.prev["b"], self.dt["b"] = t(plan.prev_b), t(plan.dt_b, torch.float32)


def _act(x, relu: bool):
    return torch.relu(x) if relu else x


def _rgcn_pre(layer, h, c: _Ctx, row0: int, row1: int, training: bool):
    """models/RGCN.py:53-104 for the packed rows [row0, row1) up to the activation: block-diagonal messages, edge norm,
    sum, node norm, (+bias), + dropout(self loop).  ``h`` holds the layer input of every packed row."""
    D = h.shape[1]
    e0, e1 = int(c.row_ptr[row0]), int(c.row_ptr[row1])
    src, rel, dst = c.e_src[e0:e1], c.e_rel[e0:e1], c.e_dst[e0:e1]
    si, so = layer.submat_in, layer.submat_out
    w = layer.weight.index_select(0, rel).view(-1, si, so)
    msg = torch.bmm(h.index_select(0, src).view(-1, 1, si), w).view(e1 - e0, D)
    msg = msg * c.norm.index_select(0, dst).unsqueeze(1)
    agg = torch.zeros(row1 - row0, D, dtype=h.dtype, device=h.device).index_add(0, dst - row0, msg)
    agg = agg * c.norm[row0:row1].unsqueeze(1)
    loop = F.dropout(h[row0:row1] @ layer.loop_weight, p=layer.dropout_p, training=training)
    return agg + (layer.h_bias if layer.bias else 0) + loop


def _iso_pre(layer, x, training: bool):
    """models/RGCN.py:78-89 up to the activation (note the residual)."""
    out = x + F.dropout(x @ layer.loop_weight, p=layer.dropout_p, training=training)
    return out + layer.h_bias if layer.bias else out


def _decay(model, layer, dt):
    if layer.learnable_lambda:                                   # models/RGCN.py:106-107
        return torch.exp(-torch.clamp(layer.exponential_decay(dt.view(-1, 1)), min=0))
    return torch.exp(-dt.view(-1, 1) * float(model.args.inv_temperature))


def _gru(model, rnn, x, h0):
    """torch.nn.GRU with sequence length 1 (models/RRGCN.py:84) or the --type1 cell
    (models/GRU_cell.py:18-31)."""
    if bool(getattr(model.args, "type1", False)):
        D = h0.shape[1]
        i_n = x @ rnn.weight_ih.t() + rnn.bias_ih
        gh = h0 @ rnn.weight_hh.t() + rnn.bias_hh
        r, z = torch.sigmoid(gh[:, :D]), torch.sigmoid(gh[:, D:2 * D])
        n = torch.tanh(i_n + r * gh[:, 2 * D:])
        return n + z * (h0 - n)
    # the nn.GRU equations written out (gate order r, z, n): plain fp32 matmuls in the forward AND the backward pass --
    # the cuDNN RNN routes both through TF32 unless torch.backends.cudnn.allow_tf32 is cleared globally
    D = h0.shape[1]
    gi = x @ rnn.weight_ih_l0.t() + rnn.bias_ih_l0
    gh = h0 @ rnn.weight_hh_l0.t() + rnn.bias_hh_l0
    r = torch.sigmoid(gi[:, :D] + gh[:, :D])
    z = torch.sigmoid(gi[:, D:2 * D] + gh[:, D:2 * D])
    n = torch.tanh(gi[:, 2 * D:] + r * gh[:, 2 * D:])
    return (1 - z) * n + z * h0


def _cell(model, layer, direction: str):
    """Recurrent parameters of one direction: RRGCN.py:75, 126; BiRRGCN.py:21-25, 109-113."""
    gru = model.args.module in ("GRRGCN", "BiGRRGCN")
    if not model.bidirectional:
        return layer.rnn if gru else layer.time_weight
    if gru:
        return layer.forward_rnn if direction == "f" else layer.backward_rnn
    return layer.time_weight_forward if direction == "f" else layer.time_weight_backward


def _recur(model, layer, pre, states, decays, dirs, relu: bool):
    """The recurrent half of one layer given its pre-activation RGCN output ``pre`` and, per direction, the previous
    state rows (zeros where there is none) and the decay arguments.  GRU flavour: RRGCN.py:77-89, BiRRGCN.py:27-81 (the
    activation applies BEFORE the cells, the directions add); linear flavour: RRGCN.py:130-154 (the recurrent terms join
    the sum before the activation)."""
    gru = model.args.module in ("GRRGCN", "BiGRRGCN")
    if gru:
        x = _act(pre, relu)
        out = 0
        for d, raw, dt in zip(dirs, states, decays):
            out = out + _gru(model, _cell(model, layer, d), x, raw * _decay(model, layer, dt))
        return out
    tot = pre
    for d, raw, dt in zip(dirs, states, decays):
        tot = tot + (raw @ _cell(model, layer, d)) * torch.exp(-dt.view(-1, 1) * float(model.args.inv_temperature))
    return _act(tot, relu)


def encode(model, plan):
    """-> (state [R, D] of every packed row, ctx, layer-1 state or None); differentiable w.r.t. the model parameters.
    Same launch structure as EncoderRuntime._build_recurrent (temp_b200/runtime.py), one torch statement per launch."""
    _check(model)
    c = _Ctx(model, plan)
    enc = model.ent_encoder
    l1, l2 = enc.layer_1, enc.layer_2
    training = model.training
    gru = model.args.module in ("GRRGCN", "BiGRRGCN")
    bi = plan.bidirectional
    relu2 = bi                                   # BiRRGCN.py:202-203 vs RRGCN.py:186-187
    use_te = enc.use_time_embedding
    D = model.embed_size
    h0 = model.ent_embeds.index_select(0, c.ent_id)
    S = torch.zeros(plan.R, D, dtype=h0.dtype, device=h0.device)
    S1 = None

    def step(layer, seg, pre, state_prev, relu):
        rows = torch.arange(seg.row0, seg.row1, device=S.device)
        dirs = ["f", "b"] if (seg.kind == "final" and bi) else [seg.kind[-1] if seg.kind != "final" else "f"]
        states, decays = [], []
        for d in dirs:
            which = "b" if (d == "b" and seg.kind == "final") else "a"
            prev = c.prev[which].index_select(0, rows)
            has = (prev >= 0).unsqueeze(1)
            states.append(torch.where(has, state_prev.index_select(0, prev.clamp(min=0)), torch.zeros((), device=S.device)))
            decays.append(c.dt[which].index_select(0, rows))
        hn = _recur(model, layer, pre, states, decays, dirs, relu)
        if use_te:
            hn = hn + layer.time_embed.index_select(0, c.row_time.index_select(0, rows))   # RRGCN.py:202-203
        return rows, hn

    if enc.rec_only_last_layer:
        h1 = _rgcn_pre(l1, h0, c, 0, plan.R, training)          # plain RGCN layer, no activation (RRGCN.py:179-185)
        x2 = _rgcn_pre(l2, h1, c, 0, plan.R, training)
        for seg in plan.segments:
            rows, hn = step(l2, seg, x2[seg.row0:seg.row1], S, relu2)
            S = S.index_copy(0, rows, hn)
    else:
        # both layers recurrent, serial in the step index; the GRU flavours feed layer 1 with the previous LAYER-2 state
        # (the reference's graph aliasing), the linear flavours keep separate states
        S1 = torch.zeros_like(S)
        for seg in plan.segments:
            rows, hn = step(l1, seg, _rgcn_pre(l1, h0, c, seg.row0, seg.row1, training), S if gru else S1, False)
            S1 = S1.index_copy(0, rows, hn)
            rows, hn = step(l2, seg, _rgcn_pre(l2, S1, c, seg.row0, seg.row1, training), S, relu2)
            S = S.index_copy(0, rows, hn)
    return S, c, S1


def all_embeds(model, plan, S, i: int, S1=None, shared=None):
    """get_all_embeds_Gt (models/DynamicRGCN.py:56-64, BiDynamicRGCN.py:102-112): forward_isolated over all entities
    with item i's history ("history forgets": only the entities of the last history step carry a state, one step old),
    rows of the target graph's entities overwritten by the graph states.

    ``shared`` (a dict kept across the items of one batch) enables the zero-history de-duplication of
    temp_b200/isolated.py for the GRU flavours with --rec-only-last-layer: an entity without history has the same
    isolated state for every item up to ``+ time_embed[t]``, so that state is evaluated once per batch for all entities
    and only the entities of the item's last history step go through the cells again with their states."""
    if bool(getattr(model.args, "use_embed_for_non_active", False)):      # DynamicRGCN.py:58-59: no isolated pass
        fin = plan.final.instances[i]
        ids = torch.as_tensor(fin.snapshot.node_ids, device=S.device).long()
        return model.ent_embeds.index_copy(0, ids, S[fin.row0:fin.row0 + fin.n])
    enc = model.ent_encoder
    l1, l2 = enc.layer_1, enc.layer_2
    training = model.training
    gru = model.args.module in ("GRRGCN", "BiGRRGCN")
    bi = plan.bidirectional
    M, D = model.num_ents, model.embed_size
    dev = S.device
    dirs = ["f", "b"] if bi else ["f"]
    lasts = [plan.last_hist_f[i]] + ([plan.last_hist_b[i]] if bi else [])
    t = int(plan.final_times[i])
    fin = plan.final.instances[i]
    fin_ids = torch.as_tensor(fin.snapshot.node_ids, device=dev).long()

    if gru and enc.rec_only_last_layer and shared is not None and not (training and l1.dropout_p > 0):
        if "base" not in shared:
            x = _act(_iso_pre(l2, _iso_pre(l1, model.ent_embeds, training), training), bi)
            zero = torch.zeros(M, D, dtype=S.dtype, device=dev)
            base = 0
            for d in dirs:
                base = base + _gru(model, _cell(model, l2, d), x, zero)
            shared["x"], shared["base"] = x, base
        x, out = shared["x"], shared["base"]
        if enc.use_time_embedding:
            out = out + l2.time_embed[t]
        sets = [torch.as_tensor(last.snapshot.node_ids, device=dev).long() for last in lasts if last is not None]
        if sets:
            ents = torch.unique(torch.cat(sets))
            val = 0
            for d, last in zip(dirs, lasts):
                h = torch.zeros(M, D, dtype=S.dtype, device=dev)
                if last is not None:
                    ids = torch.as_tensor(last.snapshot.node_ids, device=dev).long()
                    h = h.index_copy(0, ids, S[last.row0:last.row0 + last.n])
                h = h.index_select(0, ents)
                val = val + _gru(model, _cell(model, l2, d), x.index_select(0, ents),
                                 h * _decay(model, l2, torch.ones(ents.shape[0], dtype=S.dtype, device=dev)))
            if enc.use_time_embedding:
                val = val + l2.time_embed[t]
            out = out.index_copy(0, ents, val)
        return out.index_copy(0, fin_ids, S[fin.row0:fin.row0 + fin.n])

    ones = torch.ones(M, dtype=S.dtype, device=dev)

    def hist_of(state):
        out = []
        for last in lasts:
            h = torch.zeros(M, D, dtype=S.dtype, device=dev)
            if last is not None:
                ids = torch.as_tensor(last.snapshot.node_ids, device=dev).long()
                h = h.index_copy(0, ids, state[last.row0:last.row0 + last.n])
            out.append(h)
        return out

    def rec_iso(layer, x, state, relu):
        out = _recur(model, layer, _iso_pre(layer, x, training), hist_of(state), [ones] * len(dirs), dirs, relu)
        return out + layer.time_embed[t] if enc.use_time_embedding else out

    if enc.rec_only_last_layer:
        first = _iso_pre(l1, model.ent_embeds, training)
    else:
        first = rec_iso(l1, model.ent_embeds, S if gru else S1, False)
    second = rec_iso(l2, first, S, bi)
    return second.index_copy(0, fin_ids, S[fin.row0:fin.row0 + fin.n])


# ---- static and attention families (baselines/StaticRGCN.py:36-89; models/SelfAttentionRGCN.py:26-43, 73-139,
# models/BiSelfAttentionRGCN.py:17-87 over models/SARGCN.py:25-62, 103-125) ------------------------------------------------
def encode_static(model, plan):
    """models/RGCN.py:154-159 on the packed target snapshots -> [R, D]."""
    c = _Ctx(model, plan)
    enc = model.ent_encoder
    h0 = model.ent_embeds.index_select(0, c.ent_id)
    h1 = _rgcn_pre(enc.layer_1, h0, c, 0, plan.R, model.training)
    out = torch.relu(_rgcn_pre(enc.layer_2, h1, c, 0, plan.R, model.training))
    if enc.use_time_embedding:
        out = out + enc.layer_2.time_embed.index_select(0, c.row_time)
    return out


def all_embeds_static(model, plan, out, i: int):
    """baselines/StaticRGCN.py:48-58 over models/RGCN.py:161-164."""
    enc = model.ent_encoder
    t = int(plan.final_times[i])
    if bool(getattr(model.args, "use_embed_for_non_active", False)):      # StaticRGCN.py:51-52
        y = model.ent_embeds
    else:
        y = torch.relu(_iso_pre(enc.layer_2, _iso_pre(enc.layer_1, model.ent_embeds, model.training), model.training))
        if enc.use_time_embedding:
            y = y + enc.layer_2.time_embed[t]
    fin = plan.final.instances[i]
    ids = torch.as_tensor(fin.snapshot.node_ids, device=out.device).long()
    return y.index_copy(0, ids, out[fin.row0:fin.row0 + fin.n])


def _attend(layer, cur, prev, tau, mask):
    """models/SARGCN.py:25-53: cur [N, D], prev [N, T, D], additive mask [N, T + 1], slot ages tau [T + 1]; the output
    channel order is [d_k major, head minor]."""
    N, D = cur.shape
    H = layer.h
    dk = D // H
    dec = 0
    if layer.learnable_lambda:
        dec = -torch.clamp(layer.exponential_decay(tau.view(-1, 1)), min=0).view(-1)
    allv = torch.cat([prev, cur.unsqueeze(1)], dim=1)
    q = (cur @ layer.q_linear.weight.t()).view(N, 1, H, dk).transpose(1, 2)
    k = (allv @ layer.k_linear.weight.t()).view(N, -1, H, dk).transpose(1, 2)
    v = (allv @ layer.v_linear.weight.t()).view(N, -1, H, dk).transpose(1, 2)
    sc = torch.matmul(q, k.transpose(-2, -1)).view(N, H, -1) / (dk ** 0.5) + mask.unsqueeze(1) + dec
    o = torch.matmul(torch.softmax(sc, dim=-1).unsqueeze(2), v).view(N, H, dk)
    return o.transpose(1, 2).contiguous().view(N, D)


def _slot_gather(values, slots):
    """values [R, D], slots [N, T] packed rows (-1: inactive) -> ([N, T, D] with zero rows, additive mask [N, T + 1])."""
    has = slots >= 0
    prev = torch.where(has.unsqueeze(-1), values.index_select(0, slots.clamp(min=0).reshape(-1)).view(*slots.shape, -1),
                       torch.zeros((), device=values.device))
    mask = torch.where(has, 0.0, -10e9).to(values.dtype)
    return prev, torch.cat([mask, torch.zeros(slots.shape[0], 1, dtype=values.dtype, device=values.device)], dim=1)


def encode_attention(model, plan):
    """History snapshots: two plain layers, both outputs with their time embedding (SARGCN.py:103-107); final rows: attention
    over the history slots (SARGCN.py:109-117) -> (out [n_final, D], first_te [R, D], second_te [R, D], tau)."""
    c = _Ctx(model, plan)
    enc = model.ent_encoder
    l1, l2 = enc.layer_1, enc.layer_2
    dev = model.ent_embeds.device
    h0 = model.ent_embeds.index_select(0, c.ent_id)
    h1 = _rgcn_pre(l1, h0, c, 0, plan.R, model.training)
    first_te = h1 + l1.time_embed.index_select(0, c.row_time)
    second_te = torch.relu(_rgcn_pre(l2, h1, c, 0, plan.R, model.training)) + l2.time_embed.index_select(0, c.row_time)
    fin = plan.final
    tau = torch.tensor(model.time_diff(plan), dtype=torch.float32, device=dev)
    slots = torch.as_tensor(np.ascontiguousarray(plan.slot_row), device=dev).long().view(fin.row1 - fin.row0, -1)
    prev2, mask = _slot_gather(second_te, slots)
    out = _attend(l2, second_te[fin.row0:fin.row1], prev2, tau, mask)
    if not enc.rec_only_last_layer:                       # JK max over the two layers (SARGCN.py:117)
        prev1, _ = _slot_gather(first_te, slots)
        out = torch.max(_attend(l1, first_te[fin.row0:fin.row1], prev1, tau, mask), out)
    return out, first_te, second_te, tau


def all_embeds_attention(model, plan, out, first_te, second_te, tau, i: int):
    """models/SelfAttentionRGCN.py:26-43 / BiSelfAttentionRGCN.py (get_all_embeds_Gt) over SARGCN.py:55-62, 119-125."""
    enc = model.ent_encoder
    l1, l2 = enc.layer_1, enc.layer_2
    dev = out.device
    M, L = model.num_ents, plan.seq_len
    t = int(plan.final_times[i])
    if bool(getattr(model.args, "use_embed_for_non_active", False)):      # SelfAttentionRGCN.py:31-32: no isolated pass
        fin = plan.final.instances[i]
        ids = torch.as_tensor(fin.snapshot.node_ids, device=dev).long()
        return model.ent_embeds.index_copy(0, ids, out[fin.row0 - plan.final.row0:fin.row0 - plan.final.row0 + fin.n])
    slots = np.full((M, plan.n_slots), -1, dtype=np.int64)
    for k in range(L - 1):
        inst = plan.steps_f[k].get(i)
        if inst is not None:
            slots[inst.snapshot.node_ids, k] = np.arange(inst.row0, inst.row0 + inst.n)
        if plan.bidirectional:
            inst = plan.steps_b[k].get(plan.batch - 1 - i)
            if inst is not None:
                slots[inst.snapshot.node_ids, L - 1 + k] = np.arange(inst.row0, inst.row0 + inst.n)
    slots = torch.as_tensor(slots, device=dev)
    prev2, mask = _slot_gather(second_te, slots)
    if enc.rec_only_last_layer:
        first = _iso_pre(l1, model.ent_embeds, model.training)
    else:
        prev1, _ = _slot_gather(first_te, slots)
        first = _attend(l1, _iso_pre(l1, model.ent_embeds, model.training) + l1.time_embed[t], prev1, tau, mask)
    second = _attend(l2, torch.relu(_iso_pre(l2, first, model.training)) + l2.time_embed[t], prev2, tau, mask)
    table = second if enc.rec_only_last_layer else torch.max(first, second)
    fin = plan.final.instances[i]
    ids = torch.as_tensor(fin.snapshot.node_ids, device=dev).long()
    return table.index_copy(0, ids, out[fin.row0 - plan.final.row0:fin.row0 - plan.final.row0 + fin.n])


def training_loss(model, t_list):
    """models/DynamicRGCN.py:176-194 (and the Bi / attention / static twins) with autograd: sub-sampled window in train()
    mode, bit-exact negative sampling, tail + head cross-entropy through the torch scorers."""
    _check(model)
    plan = model.plan(t_list, transform=model.train_edge_sampler() if model.training else None)
    fam = model.family
    if fam == "static":
        S = encode_static(model, plan)
    elif fam == "attention":
        S, first_te, second_te, tau = encode_attention(model, plan)
    else:
        S, _, S1 = encode(model, plan)
    dev = S.device
    loss = 0
    shared = {}
    for i, (t, g) in enumerate(zip(plan.final_times, plan.final_snapshots)):
        fin = plan.final.instances[i]
        off = plan.final.row0 if fam == "attention" else 0           # (the attention output holds the final rows only)
        ent_embed = S[fin.row0 - off:fin.row0 - off + fin.n]
        triplets, neg_tail, neg_head, labels = model.corrupter.single_graph_negative_sampling(t, g, model.num_ents)
        triplets, neg_tail, neg_head, labels = (x.to(dev) for x in (triplets, neg_tail, neg_head, labels))
        if fam == "static":
            all_g = all_embeds_static(model, plan, S, i)
        elif fam == "attention":
            all_g = all_embeds_attention(model, plan, S, first_te, second_te, tau, i)
        else:
            all_g = all_embeds(model, plan, S, i, S1, shared)
        loss = loss + model.train_link_prediction(ent_embed, triplets, neg_tail, labels, all_g, corrupt_tail=True)
        loss = loss + model.train_link_prediction(ent_embed, triplets, neg_head, labels, all_g, corrupt_tail=False)
    return loss
