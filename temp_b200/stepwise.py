"""The reference's per-step ENCODER calls on the CUDA path ("the encoder object must answer
forward / forward_isolated / forward_one_direction ...").

``model.encode()`` runs a whole window batch as one launch program and is the fast entry.  Code written against the
reference's encoder objects -- its own ``DynamicRGCN`` / ``BiDynamicRGCN`` / ``StaticRGCN`` drivers, the Aggregator --
instead calls the encoder once per time step with dense tensors:

    RGCN.forward(batched_graph, time_batched_list_t, node_sizes)                              models/RGCN.py:154-159
    RGCN.forward_isolated(ent_embeds, time)                                                   models/RGCN.py:161-164
    RRGCN.forward(batched_graph, first_prev, second_prev, time_diff, times, node_sizes)       models/RRGCN.py:192-204
    RRGCN.forward_isolated(ent_embeds, first_prev, second_prev, time_diff, time)              models/RRGCN.py:206-217
    BiRRGCN.forward(batched_graph, first_prev_f, second_prev_f, dt_f, first_prev_b, second_prev_b, dt_b, times,
                    node_sizes)                                                               models/BiRRGCN.py:210-226
    BiRRGCN.forward_one_direction(batched_graph, first_prev, second_prev, dt, times, node_sizes, forward)   228-240
    BiRRGCN.forward_isolated(ent_embeds, first_prev_f, second_prev_f, dt_f, first_prev_b, second_prev_b, dt_b, time)
                                                                                              models/BiRRGCN.py:242-257
    SARGCN.forward(batched_graph, times, node_sizes)                                          models/SARGCN.py:103-107
    SARGCN.forward_final(batched_graph, prev1, prev2, time_diff, local_attn_mask, times, node_sizes)       109-117
    SARGCN.forward_isolated(ent_embeds, prev1, prev2, time_diff, local_attn_mask, time)                    119-125

Those methods (temp_b200/encoder.py) land here: one call = one short launch program of the same kernels
(``temp_rgcn_layer_fwd`` with the chained input-gate GEMM, ``temp_gru_fwd``), previous states read from the caller's
dense ``[N, D]`` tensors through an identity row map.  ``batched_graph`` is a ``BatchedSnapshots`` (the subset of
``dgl.batch`` the encoders use: ``ndata['h']``, ``ndata['id']``, ``number_of_nodes()``, ``local_var()``).
The graph-aliasing quirk of the GRU flavours (``first`` and ``second`` are the same tensor) is kept.
Inference only (no autograd); the post-ensemble / impute / EMA variants are not built.
"""
from __future__ import annotations

import weakref
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import lib
from .planner import plan_snapshots

_OWNERS = weakref.WeakKeyDictionary()          # encoder object -> weakref(model shell)


def bind(encoder, model) -> None:
    _OWNERS[encoder] = weakref.ref(model)


def owner_of(encoder):
    ref = _OWNERS.get(encoder)
    model = ref() if ref is not None else None
    if model is None:
        raise RuntimeError("temp_b200: this encoder object is not attached to a TKG_Module shell (the per-step calls run "
                           "through the shell's CUDA runtime)")
    return model


class BatchedSnapshots(object):
    """``dgl.batch(g_list)`` as the encoder calls see it (models/DynamicRGCN.py:92-93)."""

    def __init__(self, snapshots: Sequence, times: Optional[Sequence[int]] = None, ndata: Optional[dict] = None):
        self.snapshots = list(snapshots)
        self.times = [int(t) for t in (times if times is not None else [s.time for s in self.snapshots])]
        self.node_sizes = [s.num_nodes for s in self.snapshots]
        self.ndata = {} if ndata is None else ndata
        self._plan = None

    @property
    def ids(self) -> np.ndarray:
        """global entity id of every batched node"""
        return np.concatenate([s.node_ids for s in self.snapshots]) if self.snapshots else np.zeros(0, dtype=np.int64)

    def number_of_nodes(self) -> int:
        return int(sum(self.node_sizes))

    def nodes(self):
        return torch.arange(self.number_of_nodes())

    def local_var(self):
        g = BatchedSnapshots(self.snapshots, self.times, dict(self.ndata))
        g._plan = self._plan
        return g

    def plan(self, times: Optional[Sequence[int]] = None):
        times = self.times if times is None else [int(t) for t in times]
        if self._plan is None or self._plan.final_times != times:
            self._plan = plan_snapshots(self.snapshots, times)
        return self._plan


def batch(snapshots: Sequence, model=None) -> BatchedSnapshots:
    """dgl.batch + ``ndata['h'] = ent_embeds[ndata['id']]`` when a model is given (models/DynamicRGCN.py:92-93)."""
    g = BatchedSnapshots(snapshots)
    if model is not None:
        ids = torch.from_numpy(g.ids).to(model.ent_embeds.device).long()
        g.ndata["id"] = ids.view(-1, 1)
        g.ndata["h"] = model.ent_embeds.detach().index_select(0, ids)
    return g


def _times_list(times) -> List[int]:
    if torch.is_tensor(times):
        return [int(t) for t in times.reshape(-1).tolist()]
    return [int(t.item()) if torch.is_tensor(t) else int(t) for t in times]


def _dense(x, rows: int, cols: int, dev) -> torch.Tensor:
    x = torch.as_tensor(x, device=dev).detach().to(torch.float32).reshape(rows, cols).contiguous()
    return x


class _Step(object):
    """Shared scaffolding of one encoder call."""

    def __init__(self, encoder):
        self.enc = encoder
        self.m = m = owner_of(encoder)
        self.rt = m.runtime                                   # raises without a CUDA device: no CPU route
        self.dev = m.ent_embeds.device
        self.D = m.embed_size
        self.gru = m.args.module in ("GRRGCN", "BiGRRGCN")
        self.bi = m.bidirectional
        self.type1 = bool(getattr(m.args, "type1", False))
        self.G = self.D if self.type1 else 3 * self.D
        self.prog = lib.Program()

    def run(self):
        self.prog.run()
        self.rt.mark_run("step")                              # the pinned staging buffer is reused call to call

    def cell(self, layer, d):
        if not self.bi:
            return ("rnn", layer.rnn) if self.gru else ("time_weight", layer.time_weight)
        if self.gru:
            return ("forward_rnn", layer.forward_rnn) if d == "f" else ("backward_rnn", layer.backward_rnn)
        return (("time_weight_forward", layer.time_weight_forward) if d == "f"
                else ("time_weight_backward", layer.time_weight_backward))

    def recurrent(self, layer, lname, rows, dirs, prevs, dts, ident, make_layer, out, relu, te, gru_kw, local=None,
                  local_te=False, gi_override=None):
        """One recurrent layer over ``rows``: ``make_layer(terms=..., **outputs)`` builds its RGCN / isolated half.
        ``local``: also store the layer's pre-recurrence output (the post-ensemble "local" stream, RRGCN.py:227-233), with the
        time embedding when ``local_te``.  ``gi_override``: input pre-activations computed by the caller (impute)."""
        rt, n = self.rt, rows[1] - rows[0]
        if self.gru:
            cells = [self.cell(layer, d) for d in dirs]
            w, b = rt._wih(lname, cells)
            GL = self.G * len(dirs)
            if gi_override is not None:
                gi = gi_override
            else:
                gi = torch.empty(max(n, 1), GL, dtype=torch.float32, device=self.dev)
                self.prog.keepalive.append(gi)
                outputs = dict(chain=(w, b, gi, GL))
                if local is not None:
                    outputs.update(h_out=local, te_out=local_te)
                self.prog.add(lib.OP_LAYER, make_layer(layer, relu, [], **outputs))
            for j, d in enumerate(dirs):
                self.prog.add(lib.OP_GRU, rt._gru(layer, cells[j][1], cells[j][0], rows, gi=gi, gi_ld=GL, gi_off=j * self.G,
                                                  state=prevs[j], prev=ident.data_ptr(), dt=dts[j].data_ptr(), out=out,
                                                  te=te and j == len(dirs) - 1, accumulate=j > 0, layer_name=lname,
                                                  **gru_kw))
        else:
            extra = [rt._term(prevs[j], self.cell(layer, d)[1], index=ident.data_ptr(), dt=dts[j].data_ptr())
                     for j, d in enumerate(dirs)]
            self.prog.add(lib.OP_LAYER, make_layer(layer, relu, extra, h_out=out, te_out=te))


def graph_step(encoder, bg: BatchedSnapshots, times, prev1, prev2, dts, dirs, want_local=False):
    """One encoder call on a batch of snapshots.  prev1 / prev2 / dts: per direction in ``dirs`` the dense previous
    states of layer 1 / layer 2 ``[N, D]`` and the time differences ``[N]`` (prev1 may be None with
    --rec-only-last-layer).  -> (first, second), or (local, first, second) with ``want_local`` (the post-ensemble calls:
    local = the layer-2 RGCN output before the cells, + its time embedding; GRU flavours only)."""
    st = _Step(encoder)
    m, rt, D, dev = st.m, st.rt, st.D, st.dev
    if m.family != "recurrent":
        raise NotImplementedError("graph_step serves the recurrent encoders")
    plan = bg.plan(_times_list(times))
    R = plan.R
    rows = (0, R)
    x = _dense(bg.ndata["h"], R, D, dev)
    first = torch.empty(R, D, dtype=torch.float32, device=dev)
    second = torch.empty(R, D, dtype=torch.float32, device=dev)
    local = torch.empty(R, D, dtype=torch.float32, device=dev) if want_local else None
    if want_local and not st.gru:
        raise NotImplementedError("the post-ensemble calls exist for the GRU flavours only (models/RRGCN.py:219-234)")
    if R == 0:
        return (local, first, second) if want_local else (first, second)
    dptr = rt.stage_plan(plan, st.prog, tag="step")
    ident = torch.arange(R, dtype=torch.int32, device=dev)
    prev2 = [_dense(p, R, D, dev) for p in prev2]
    dts = [_dense(t, R, 1, dev).view(-1) for t in dts]
    st.prog.keepalive += [x, ident, first, second] + prev2 + dts
    l1, l2 = encoder.layer_1, encoder.layer_2
    use_te = encoder.use_time_embedding
    relu2 = st.bi                                             # BiRRGCN.py:202-203 vs RRGCN.py:186-187

    def graph_layer(x_in):
        def make(layer, relu, extra, **outputs):
            return rt._layer(layer, rows, dptr, x=x_in, x_is_embed=False, act=relu,
                             terms=[rt._term(x_in, layer.loop_weight)] + extra, **outputs)
        return make

    if encoder.rec_only_last_layer:
        st.prog.add(lib.OP_LAYER, graph_layer(x)(l1, False, [], h_out=first))
    else:
        prev1 = [_dense(p, R, D, dev) for p in prev1]
        st.prog.keepalive += prev1
        st.recurrent(l1, "layer_1", rows, dirs, prev1, dts, ident, graph_layer(x), first, False, use_te, dict(dptr=dptr))
    st.recurrent(l2, "layer_2", rows, dirs, prev2, dts, ident, graph_layer(first), second, relu2, use_te, dict(dptr=dptr),
                 local=local, local_te=use_te)
    if local is not None:
        st.prog.keepalive.append(local)
    st.run()
    if st.gru:
        first = second                                        # the layer-2 GRU writes into the shared graph
    return (local, first, second) if want_local else (first, second)


def impute_weights(encoder, dts, dirs):
    """RRGCN.calc_impute_weight (models/RRGCN.py:271-272); the Bi model halves each direction's weight
    (models/BiRRGCN.py:309-310, 331-332).  dts: per direction ``[M]`` -> per direction ``[M, 1]``."""
    if len(dirs) == 1:
        lin = encoder.impute_weight
        return [torch.exp(-torch.clamp(dts[0].view(-1, 1) * lin.weight.detach().view(()) + lin.bias.detach().view(()), min=0))]
    out = []
    for dt, lin in zip(dts, (encoder.impute_weight_forward, encoder.impute_weight_backward)):
        out.append(torch.exp(-torch.clamp(dt.view(-1, 1) * lin.weight.detach().view(()) + lin.bias.detach().view(()), min=0)) / 2)
    return out


def _blend(ws, locs, x):
    acc, rest = 0, 1
    for w, loc in zip(ws, locs):
        acc = acc + w * loc
        rest = rest - w
    return acc + rest * x


def isolated_step(encoder, ent_embeds, t: int, prev1, prev2, dts, dirs, mode="plain", locs=None):
    """``forward_isolated`` over all rows of ``ent_embeds`` (RRGCN.py:206-217, BiRRGCN.py:242-257) -> second [M, D].

    mode 'post': ``forward_post_ensemble_isolated`` (RRGCN.py:236-254, BiRRGCN.py:295-319) -> (local, second): the layer
        launch also stores the layer-2 output before the cells; with an impute encoder that local stream is blended with the
        history's local rows ``locs`` (element-wise, torch) before the time embedding is added.
    mode 'impute': ``forward_isolated_impute`` (RRGCN.py:256-269, BiRRGCN.py:321-338): the layer-2 cells read the BLENDED
        stream, so their input pre-activations are one library GEMM on the blended rows (torch.addmm) between the layer
        launch and the recurrent launch -- a "next" row, not the hot path."""
    st = _Step(encoder)
    m, rt, D, dev = st.m, st.rt, st.D, st.dev
    if m.family != "recurrent":
        raise NotImplementedError("isolated_step serves the recurrent encoders")
    M = int(ent_embeds.shape[0])
    rows = (0, M)
    t = int(t.item()) if torch.is_tensor(t) else int(t)
    x = _dense(ent_embeds, M, D, dev)
    first = torch.empty(M, D, dtype=torch.float32, device=dev)
    second = torch.empty(M, D, dtype=torch.float32, device=dev)
    ident = torch.arange(M, dtype=torch.int32, device=dev)
    prev2 = [_dense(p, M, D, dev) for p in prev2]
    dts = [_dense(v, M, 1, dev).view(-1) for v in dts]
    st.prog.keepalive += [x, ident, first, second] + prev2 + dts
    rt._live = st.prog.keepalive
    l1, l2 = encoder.layer_1, encoder.layer_2
    use_te = encoder.use_time_embedding

    def iso_layer(x_in):
        def make(layer, relu, extra, **outputs):
            return rt._layer(layer, rows, None, x=None, x_is_embed=False, graph=False, residual=True, act=relu,
                             terms=[rt._term(x_in, layer.loop_weight)] + extra, row_time_scalar=t, **outputs)
        return make

    if encoder.rec_only_last_layer:
        st.prog.add(lib.OP_LAYER, iso_layer(x)(l1, False, [], h_out=first))
    else:
        prev1 = [_dense(p, M, D, dev) for p in prev1]
        st.prog.keepalive += prev1
        st.recurrent(l1, "layer_1", rows, dirs, prev1, dts, ident, iso_layer(x), first, False, use_te,
                     dict(row_time_scalar=t))
    if mode == "plain":
        st.recurrent(l2, "layer_2", rows, dirs, prev2, dts, ident, iso_layer(first), second, st.bi, use_te,
                     dict(row_time_scalar=t))
        st.run()
        return second
    if not st.gru:
        raise NotImplementedError("the post-ensemble / impute calls exist for the GRU flavours only (models/RRGCN.py:219-272)")
    local = torch.empty(M, D, dtype=torch.float32, device=dev)
    st.prog.keepalive.append(local)
    te2 = l2.time_embed.detach()[t] if use_te else None
    if mode == "post":
        st.recurrent(l2, "layer_2", rows, dirs, prev2, dts, ident, iso_layer(first), second, st.bi, use_te,
                     dict(row_time_scalar=t), local=local, local_te=False)
        st.run()
        if encoder.impute:
            local = _blend(impute_weights(encoder, dts, dirs), [_dense(x_, M, D, dev) for x_ in locs], local)
        return (local + te2 if use_te else local), second
    # impute: layer launch (no chained GEMM) -> blend -> gi by a library GEMM -> the cells
    st.prog.add(lib.OP_LAYER, iso_layer(first)(l2, st.bi, [], h_out=local))
    st.run()
    x = _blend(impute_weights(encoder, dts, dirs), [_dense(x_, M, D, dev) for x_ in locs], local)
    st2 = _Step(encoder)
    cells = [st2.cell(l2, d) for d in dirs]
    w, b = rt._wih("layer_2", cells)
    gi = torch.addmm(b, x, w)
    st2.prog.keepalive += [x, gi, ident, second] + prev2 + dts
    st2.recurrent(l2, "layer_2", rows, dirs, prev2, dts, ident, None, second, st.bi, use_te, dict(row_time_scalar=t),
                  gi_override=gi)
    st2.run()
    return second


def static_graph_step(encoder, bg: BatchedSnapshots, times):
    """RGCN.forward (models/RGCN.py:154-159): returns the batched graph with ``ndata['h']`` = layer-2 output (+ te)."""
    st = _Step(encoder)
    m, rt, D, dev = st.m, st.rt, st.D, st.dev
    plan = bg.plan(_times_list(times))
    R = plan.R
    rows = (0, R)
    out_g = bg.local_var()
    x = _dense(bg.ndata["h"], R, D, dev)
    h1 = torch.empty(R, D, dtype=torch.float32, device=dev)
    out = torch.empty(R, D, dtype=torch.float32, device=dev)
    if R > 0:
        dptr = rt.stage_plan(plan, st.prog, tag="step")
        st.prog.keepalive += [x, h1, out]
        l1, l2 = encoder.layer_1, encoder.layer_2
        st.prog.add(lib.OP_LAYER, rt._layer(l1, rows, dptr, x=x, x_is_embed=False, act=False,
                                            terms=[rt._term(x, l1.loop_weight)], h_out=h1))
        st.prog.add(lib.OP_LAYER, rt._layer(l2, rows, dptr, x=h1, x_is_embed=False, act=True,
                                            terms=[rt._term(h1, l2.loop_weight)], h_out=out,
                                            te_out=encoder.use_time_embedding))
        st.run()
    out_g.ndata["h"] = out
    return out_g


def static_isolated_step(encoder, ent_embeds, t):
    """RGCN.forward_isolated (models/RGCN.py:161-164)."""
    st = _Step(encoder)
    rt, D, dev = st.rt, st.D, st.dev
    M = int(ent_embeds.shape[0])
    rows = (0, M)
    t = int(t.item()) if torch.is_tensor(t) else int(t)
    x = _dense(ent_embeds, M, D, dev)
    y1 = torch.empty(M, D, dtype=torch.float32, device=dev)
    out = torch.empty(M, D, dtype=torch.float32, device=dev)
    st.prog.keepalive += [x, y1, out]
    rt._live = st.prog.keepalive
    l1, l2 = encoder.layer_1, encoder.layer_2
    kw = dict(x=None, x_is_embed=False, graph=False, residual=True)
    st.prog.add(lib.OP_LAYER, rt._layer(l1, rows, None, act=False, terms=[rt._term(x, l1.loop_weight)], h_out=y1, **kw))
    st.prog.add(lib.OP_LAYER, rt._layer(l2, rows, None, act=True, terms=[rt._term(y1, l2.loop_weight)], h_out=out,
                                        te_out=encoder.use_time_embedding, row_time_scalar=t, **kw))
    st.run()
    return out


# ---- SARGCN (models/SARGCN.py:103-125) -----------------------------------------------------------------------------
def _attn_inputs(st, prevs, mask, time_diff, n):
    """Dense ``prev [n, T, D]`` histories + additive ``mask [n, T + 1]`` (0 = active slot, -10e9 = inactive;
    models/SelfAttentionRGCN.py:104-120) -> flattened history rows, slot map (-1 = inactive), tau."""
    D, dev = st.D, st.dev
    prevs = [None if p is None else torch.as_tensor(p, device=dev).detach().float().contiguous() for p in prevs]
    T = int(next(p for p in prevs if p is not None).shape[1])
    mask = torch.as_tensor(mask, device=dev).detach().float().reshape(n, T + 1)
    rows = torch.arange(n * T, dtype=torch.int32, device=dev).view(n, T)
    slot = torch.where(mask[:, :T] > -1.0, rows, torch.full_like(rows, -1)).contiguous()
    tau = torch.as_tensor(time_diff, device=dev).detach().float().reshape(-1).contiguous()
    if tau.numel() != T + 1:
        raise ValueError("time_diff must have one entry per history slot plus the current step")
    flat = [None if p is None else p.reshape(n * T, D) for p in prevs]
    st.prog.keepalive += [slot, tau] + [f for f in flat if f is not None]
    return flat, slot, tau, T


def _attn_layer(st, layer, lname, make_layer, hist_flat, slot, tau, T, n, out, combine, h_out=None, h_out_te=False):
    """q, k, v of the current rows by the chained GEMM of the layer launch (time embedding added on the way in), k, v of
    the dense history rows by one library GEMM, then the attention kernel (SARGCN.py:25-53)."""
    rt, D, dev = st.rt, st.D, st.dev
    kvw = [layer.k_linear.weight, layer.v_linear.weight]
    qkvw = [layer.q_linear.weight] + kvw
    kv_t = rt.prep.cat_t(lname + ".kv_t", kvw)
    qkv_t = rt.prep.cat_t(lname + ".qkv_t", qkvw)
    if rt.use_tc:
        rt._chain_packed[qkv_t.data_ptr()] = rt.prep.packed(lname + ".qkv_packed", qkv_t, qkvw)
    qkv = torch.empty(max(n, 1), 3 * D, dtype=torch.float32, device=dev)
    kv_hist = torch.mm(hist_flat, kv_t) if T > 0 else None
    st.prog.keepalive += [qkv, kv_hist]
    kw = dict(te_chain=True, chain=(qkv_t, None, qkv, 3 * D))
    if h_out is not None:
        kw["h_out"] = h_out
        kw["te_out"] = h_out_te
    st.prog.add(lib.OP_LAYER, make_layer(layer, **kw))
    a = lib.AttnArgs()
    a.row0, a.row1, a.d, a.heads = 0, n, D, layer.h
    a.qkv = qkv.data_ptr()
    a.kv_hist = kv_hist.data_ptr() if T > 0 else None
    a.slot_row = slot.data_ptr() if T > 0 else None
    a.n_slots = T
    a.tau = tau.data_ptr()
    wb = rt._decay_wb(layer, lname)
    a.decay_wb = None if wb is None else wb.data_ptr()
    a.combine_max = int(combine)
    a.out = out.data_ptr()
    st.prog.add(lib.OP_ATTN, a)


def attention_history_step(encoder, bg: BatchedSnapshots, times):
    """SARGCN.forward (models/SARGCN.py:103-107): two plain layers; both outputs carry their time embedding, layer 2
    consumes layer 1 WITHOUT it.  -> (first + te1, second + te2)."""
    st = _Step(encoder)
    rt, D, dev = st.rt, st.D, st.dev
    plan = bg.plan(_times_list(times))
    R = plan.R
    rows = (0, R)
    x = _dense(bg.ndata["h"], R, D, dev)
    h1 = torch.empty(R, D, dtype=torch.float32, device=dev)
    first = torch.empty(R, D, dtype=torch.float32, device=dev)
    second = torch.empty(R, D, dtype=torch.float32, device=dev)
    if R == 0:
        return first, second
    dptr = rt.stage_plan(plan, st.prog, tag="step")
    st.prog.keepalive += [x, h1, first, second]
    l1, l2 = encoder.layer_1, encoder.layer_2
    st.prog.add(lib.OP_LAYER, rt._layer(l1, rows, dptr, x=x, x_is_embed=False, act=False,
                                        terms=[rt._term(x, l1.loop_weight)], h_out=h1))
    st.prog.add(lib.OP_LAYER, rt._layer(l2, rows, dptr, x=h1, x_is_embed=False, act=True,
                                        terms=[rt._term(h1, l2.loop_weight)], h_out=second, te_out=True))
    for inst in plan.final.instances:                       # first = h1 + te1[t_g], one row-block copy per graph
        a = lib.ScatterArgs()
        a.n, a.d = int(inst.n), D
        a.src = h1.data_ptr() + inst.row0 * D * 4
        a.dst = first.data_ptr() + inst.row0 * D * 4
        a.src_index = a.dst_index = None
        a.add_row = l1.time_embed.data_ptr() + int(inst.time) * D * 4
        st.prog.add(lib.OP_SCATTER, a)
    st.run()
    return first, second


def attention_final_step(encoder, bg: BatchedSnapshots, prev1, prev2, time_diff, mask, times):
    """SARGCN.forward_final (models/SARGCN.py:109-117) -> attention output of the batched nodes."""
    st = _Step(encoder)
    rt, D, dev = st.rt, st.D, st.dev
    plan = bg.plan(_times_list(times))
    R = plan.R
    rows = (0, R)
    out = torch.empty(R, D, dtype=torch.float32, device=dev)
    if R == 0:
        return out
    x = _dense(bg.ndata["h"], R, D, dev)
    h1 = torch.empty(R, D, dtype=torch.float32, device=dev)
    dptr = rt.stage_plan(plan, st.prog, tag="step")
    (f1, f2), slot, tau, T = _attn_inputs(st, [None if encoder.rec_only_last_layer else prev1, prev2], mask, time_diff, R)
    st.prog.keepalive += [x, h1, out]
    l1, l2 = encoder.layer_1, encoder.layer_2

    def graph_layer(x_in, act):
        def make(layer, **outputs):
            return rt._layer(layer, rows, dptr, x=x_in, x_is_embed=False, act=act,
                             terms=[rt._term(x_in, layer.loop_weight)], **outputs)
        return make

    if encoder.rec_only_last_layer:
        st.prog.add(lib.OP_LAYER, graph_layer(x, False)(l1, h_out=h1))
    else:
        _attn_layer(st, l1, "layer_1", graph_layer(x, False), f1, slot, tau, T, R, out, False, h_out=h1)
    _attn_layer(st, l2, "layer_2", graph_layer(h1, True), f2, slot, tau, T, R, out, not encoder.rec_only_last_layer)   # JK max
    st.run()
    return out


def attention_isolated_step(encoder, ent_embeds, prev1, prev2, time_diff, mask, t):
    """SARGCN.forward_isolated (models/SARGCN.py:119-125)."""
    st = _Step(encoder)
    rt, D, dev = st.rt, st.D, st.dev
    M = int(ent_embeds.shape[0])
    rows = (0, M)
    t = int(t.item()) if torch.is_tensor(t) else int(t)
    x = _dense(ent_embeds, M, D, dev)
    first = torch.empty(M, D, dtype=torch.float32, device=dev)
    out = torch.empty(M, D, dtype=torch.float32, device=dev)
    (f1, f2), slot, tau, T = _attn_inputs(st, [None if encoder.rec_only_last_layer else prev1, prev2], mask, time_diff, M)
    st.prog.keepalive += [x, first, out]
    rt._live = st.prog.keepalive
    l1, l2 = encoder.layer_1, encoder.layer_2

    def iso_layer(x_in, act):
        def make(layer, **outputs):
            return rt._layer(layer, rows, None, x=None, x_is_embed=False, graph=False, residual=True, act=act,
                             terms=[rt._term(x_in, layer.loop_weight)], row_time_scalar=t, **outputs)
        return make

    if encoder.rec_only_last_layer:
        st.prog.add(lib.OP_LAYER, iso_layer(x, False)(l1, h_out=first))
        _attn_layer(st, l2, "layer_2", iso_layer(first, True), f2, slot, tau, T, M, out, False)
    else:
        _attn_layer(st, l1, "layer_1", iso_layer(x, False), f1, slot, tau, T, M, first, False)
        a = lib.ScatterArgs()                                # out = first, then max with the layer-2 attention (SARGCN.py:125)
        a.n, a.d, a.src, a.dst = M, D, first.data_ptr(), out.data_ptr()
        a.src_index = a.dst_index = a.add_row = None
        st.prog.add(lib.OP_SCATTER, a)
        _attn_layer(st, l2, "layer_2", iso_layer(first, True), f2, slot, tau, T, M, out, True)
    st.run()
    return out


def attention_post_ensemble_step(encoder, bg: BatchedSnapshots, prev2, time_diff, mask, times):
    """SARGCN.forward_post_ensemble (models/SARGCN.py:137-141, the --post-aggregation layer return of SARGCN.py:44-45):
    plain layer 1, then layer 2 with its attention -> (second_local = layer-2 output + time embedding, attention output)."""
    st = _Step(encoder)
    rt, D, dev = st.rt, st.D, st.dev
    plan = bg.plan(_times_list(times))
    R = plan.R
    rows = (0, R)
    out = torch.empty(R, D, dtype=torch.float32, device=dev)
    local = torch.empty(R, D, dtype=torch.float32, device=dev)
    if R == 0:
        return local, out
    x = _dense(bg.ndata["h"], R, D, dev)
    h1 = torch.empty(R, D, dtype=torch.float32, device=dev)
    dptr = rt.stage_plan(plan, st.prog, tag="step")
    (f2,), slot, tau, T = _attn_inputs(st, [prev2], mask, time_diff, R)
    st.prog.keepalive += [x, h1, out, local]
    l1, l2 = encoder.layer_1, encoder.layer_2
    st.prog.add(lib.OP_LAYER, rt._layer(l1, rows, dptr, x=x, x_is_embed=False, act=False,
                                        terms=[rt._term(x, l1.loop_weight)], h_out=h1))

    def make(layer, **outputs):
        return rt._layer(layer, rows, dptr, x=h1, x_is_embed=False, act=True, terms=[rt._term(h1, layer.loop_weight)], **outputs)
    _attn_layer(st, l2, "layer_2", make, f2, slot, tau, T, R, out, False, h_out=local, h_out_te=True)
    st.run()
    return local, out


def attention_isolated_post_ensemble_step(encoder, ent_embeds, prev2, time_diff, mask, t):
    """SARGCN.forward_isolated_post_ensemble (models/SARGCN.py:143-146) -> (second_local, attention output)."""
    st = _Step(encoder)
    rt, D, dev = st.rt, st.D, st.dev
    M = int(ent_embeds.shape[0])
    rows = (0, M)
    t = int(t.item()) if torch.is_tensor(t) else int(t)
    x = _dense(ent_embeds, M, D, dev)
    first = torch.empty(M, D, dtype=torch.float32, device=dev)
    out = torch.empty(M, D, dtype=torch.float32, device=dev)
    local = torch.empty(M, D, dtype=torch.float32, device=dev)
    (f2,), slot, tau, T = _attn_inputs(st, [prev2], mask, time_diff, M)
    st.prog.keepalive += [x, first, out, local]
    rt._live = st.prog.keepalive
    l1, l2 = encoder.layer_1, encoder.layer_2

    def iso(x_in, act):
        def make(layer, **outputs):
            return rt._layer(layer, rows, None, x=None, x_is_embed=False, graph=False, residual=True, act=act,
                             terms=[rt._term(x_in, layer.loop_weight)], row_time_scalar=t, **outputs)
        return make
    st.prog.add(lib.OP_LAYER, iso(x, False)(l1, h_out=first))
    _attn_layer(st, l2, "layer_2", iso(first, True), f2, slot, tau, T, M, out, False, h_out=local, h_out_te=True)
    st.run()
    return local, out
