"""All-entity table ``get_all_embeds_Gt`` (region R2) on the CUDA path.

Reference: ``forward_isolated`` over ALL ``num_ents`` rows with the batch item's dense history, then
the rows of active entities overwritten one by one in a python loop
(models/DynamicRGCN.py:56-64, BiDynamicRGCN.py:102-112, SelfAttentionRGCN.py:26-43,
baselines/StaticRGCN.py:48-58; encoders RRGCN.py:206-217, BiRRGCN.py:242-257, SARGCN.py:119-125,
RGCN.py:161-164).

Here: the dense history is replaced by an int32 ``entity -> packed row`` map per item
(``prev_iso``), the isolated layers run through the same fused layer kernel (no graph part,
``residual`` = the ``x +`` of RGCN.py:83) and the overwrite is one scatter launch.

Exact algebraic saving for the GRU flavours with --rec-only-last-layer (the benchmark config):
an entity without history has the same isolated state for every batch item up to ``+ time_embed[t]``,
so the zero-history GRU (whose recurrent GEMM vanishes) is evaluated ONCE per forward for all
entities, and only the <= N_{L-2} entities that were active at the last history step run the
recurrent half per item.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import lib
from .runtime import EncodeResult, _F32


def _iso_maps(model, res: EncodeResult):
    """Per item / direction: entity -> packed row of the last history step (-1 = no state), uploaded once."""
    cache = getattr(res, "_iso", None)
    if cache is not None:
        return cache
    plan, M = res.plan, model.num_ents
    dev = model.ent_embeds.device
    dirs = ["f", "b"] if plan.bidirectional else ["f"]
    maps = np.full((len(dirs), plan.batch, M), -1, dtype=np.int32)
    for d, lasts in enumerate([plan.last_hist_f, plan.last_hist_b][:len(dirs)]):
        for i, inst in enumerate(lasts):
            if inst is not None:
                maps[d, i, inst.snapshot.node_ids] = np.arange(inst.row0, inst.row0 + inst.n, dtype=np.int32)
    cache = {"prev": torch.from_numpy(maps).to(dev), "ones": torch.ones(M, dtype=torch.float32, device=dev)}
    if model.family == "attention":
        slots = np.full((plan.batch, M, plan.n_slots), -1, dtype=np.int32)
        L = plan.seq_len
        for i in range(plan.batch):
            for k in range(L - 1):
                inst = plan.steps_f[k].get(i)
                if inst is not None:
                    slots[i, inst.snapshot.node_ids, k] = np.arange(inst.row0, inst.row0 + inst.n, dtype=np.int32)
                if plan.bidirectional:
                    inst = plan.steps_b[k].get(plan.batch - 1 - i)
                    if inst is not None:
                        slots[i, inst.snapshot.node_ids, L - 1 + k] = np.arange(inst.row0, inst.row0 + inst.n, dtype=np.int32)
        cache["slots"] = torch.from_numpy(slots).to(dev)
    res._iso = cache
    return cache


def _scatter(src, dst, n, d, src_index=None, dst_index=None, add_row=None):
    a = lib.ScatterArgs()
    a.n, a.d = int(n), int(d)
    a.src, a.dst = src, dst
    a.src_index, a.dst_index, a.add_row = src_index, dst_index, add_row
    return a


def all_embeds_item(model, res: EncodeResult, i: int, hist_item=None) -> torch.Tensor:
    """All-entity table of target graph ``i``.  ``hist_item`` (default ``i``): the batch item whose history feeds the
    isolated pass -- the reference's calc_metrics uses a lagging index there after an evaluation graph without edges
    (models/DynamicRGCN.py:196-220: ``continue`` before ``i += 1``); the overwritten rows and the time embedding are
    always graph i's."""
    rt = model.runtime
    h = i if hist_item is None else int(hist_item)
    plan, D, M = res.plan, model.embed_size, model.num_ents
    enc = model.ent_encoder
    l1, l2 = enc.layer_1, enc.layer_2
    E = model.ent_embeds
    t = plan.final_times[i]
    dev = E.device
    prog = lib.Program()
    out = torch.empty(M, D, dtype=torch.float32, device=dev)
    rows = (0, M)
    ws = rt.ws
    y1 = ws.get("iso_y1", M * D)[:M * D].view(M, D)
    fam = model.family
    use_te = enc.use_time_embedding

    def iso_layer(layer, x, act, extra_terms=(), **kw):
        return rt._layer(layer, rows, None, x=None, x_is_embed=False, graph=False, residual=True, act=act,
                         terms=[rt._term(x, layer.loop_weight)] + list(extra_terms), **kw)

    if bool(getattr(model.args, "use_embed_for_non_active", False)):
        # --use-embed-for-non-active: entities without an edge at t keep their input embedding, no isolated pass
        # (models/DynamicRGCN.py:58-59, BiDynamicRGCN.py:105-106, SelfAttentionRGCN.py:31-32, baselines/StaticRGCN.py:51-52)
        prog.add(lib.OP_SCATTER, _scatter(E.data_ptr(), out.data_ptr(), M, D))
    elif fam == "static":
        prog.add(lib.OP_LAYER, iso_layer(l1, E, False, h_out=y1))
        prog.add(lib.OP_LAYER, iso_layer(l2, y1, True, h_out=out, te_out=use_te, row_time_scalar=t))
    elif fam == "attention":
        maps = _iso_maps(model, res)
        slot_ptr = maps["slots"][h].data_ptr()
        tau = torch.tensor(model.time_diff(plan), dtype=torch.float32, device=dev)
        prog.keepalive.append(tau)
        qkv = ws.get("iso_qkv", M * 3 * D)[:M * 3 * D].view(M, 3 * D)

        def attend(layer, lname, kv_hist, dst, combine):
            a = lib.AttnArgs()
            a.row0, a.row1, a.d, a.heads = 0, M, D, layer.h
            a.qkv = qkv.data_ptr()
            a.kv_hist = kv_hist.data_ptr() if plan.hist_rows > 0 else None
            a.slot_row = slot_ptr if plan.hist_rows > 0 and plan.n_slots > 0 else None
            a.n_slots = plan.n_slots if plan.hist_rows > 0 else 0
            a.tau = tau.data_ptr()
            wb = rt._decay_wb(layer, lname)
            a.decay_wb = None if wb is None else wb.data_ptr()
            a.combine_max = int(combine)
            a.out = dst.data_ptr()
            prog.add(lib.OP_ATTN, a)

        if enc.rec_only_last_layer:
            prog.add(lib.OP_LAYER, iso_layer(l1, E, False, h_out=y1))
            first = y1
        else:
            w = rt.prep.cat_t("layer_1.qkv_t", [l1.q_linear.weight, l1.k_linear.weight, l1.v_linear.weight])
            prog.add(lib.OP_LAYER, iso_layer(l1, E, False, te_chain=True, row_time_scalar=t, chain=(w, None, qkv, 3 * D)))
            first = ws.get("iso_first", M * D)[:M * D].view(M, D)
            attend(l1, "layer_1", res.bufs["kv1"] if "kv1" in res.bufs else res.bufs["kv2"], first, False)
        w = rt.prep.cat_t("layer_2.qkv_t", [l2.q_linear.weight, l2.k_linear.weight, l2.v_linear.weight])
        prog.add(lib.OP_LAYER, iso_layer(l2, first, True, te_chain=True, row_time_scalar=t, chain=(w, None, qkv, 3 * D)))
        if enc.rec_only_last_layer:
            attend(l2, "layer_2", res.bufs["kv2"], out, False)
        else:   # max(first, second), SARGCN.py:125
            prog.add(lib.OP_SCATTER, _scatter(first.data_ptr(), out.data_ptr(), M, D))
            attend(l2, "layer_2", res.bufs["kv2"], out, True)
    else:
        maps = _iso_maps(model, res)
        gru = model.args.module in ("GRRGCN", "BiGRRGCN")
        bi = plan.bidirectional
        dirs = ["f", "b"] if bi else ["f"]
        prev = [maps["prev"][d][h].data_ptr() for d in range(len(dirs))]
        ones = maps["ones"].data_ptr()
        type1 = bool(getattr(model.args, "type1", False))
        G = D if type1 else 3 * D
        S = res.state
        relu2 = bi

        def rnn_of(layer, d):
            if not bi:
                return ("rnn", layer.rnn) if gru else ("time_weight", layer.time_weight)
            if gru:
                return ("forward_rnn", layer.forward_rnn) if d == "f" else ("backward_rnn", layer.backward_rnn)
            return (("time_weight_forward", layer.time_weight_forward) if d == "f"
                    else ("time_weight_backward", layer.time_weight_backward))

        def rec_iso(layer, lname, x, state_prev, dst, relu, te):
            if gru:
                rnns = [rnn_of(layer, d) for d in dirs]
                w, b = rt._wih(lname, rnns)
                gi = ws.get("iso_gi", M * 2 * G)[:M * 2 * G].view(M, 2 * G)
                prog.add(lib.OP_LAYER, iso_layer(layer, x, relu, chain=(w, b, gi, 2 * G)))
                for j, d in enumerate(dirs):
                    prog.add(lib.OP_GRU, rt._gru(layer, rnns[j][1], rnns[j][0], rows, gi=gi, gi_ld=2 * G, gi_off=j * G,
                                                 state=state_prev, prev=prev[j], dt=ones, out=dst,
                                                 te=te and j == len(dirs) - 1, accumulate=j > 0, row_time_scalar=t,
                                                 layer_name=lname))
            else:
                extra = [rt._term(state_prev, rnn_of(layer, d)[1], index=prev[j], dt=ones) for j, d in enumerate(dirs)]
                prog.add(lib.OP_LAYER, iso_layer(layer, x, relu, extra_terms=extra, h_out=dst, te_out=te,
                                                 row_time_scalar=t))

        if gru and enc.rec_only_last_layer:
            # ---- zero-history de-duplication (see module docstring) --------------------------------
            rnns = [rnn_of(l2, d) for d in dirs]
            base = getattr(res, "_iso_base", None)
            if base is None:
                w, b = rt._wih("layer_2", rnns)
                gi = ws.get("iso_gi", M * 2 * G)[:M * 2 * G].view(M, 2 * G)
                bs = ws.get("iso_base", M * D)[:M * D].view(M, D)
                prog.add(lib.OP_LAYER, iso_layer(l1, E, False, h_out=y1))
                prog.add(lib.OP_LAYER, iso_layer(l2, y1, relu2, chain=(w, b, gi, 2 * G)))
                for j, d in enumerate(dirs):
                    prog.add(lib.OP_GRU, rt._gru(l2, rnns[j][1], rnns[j][0], rows, gi=gi, gi_ld=2 * G, gi_off=j * G,
                                                 state=None, prev=None, dt=None, out=bs, te=False, accumulate=j > 0,
                                                 layer_name="layer_2"))
                base = res._iso_base = (bs, gi)
            bs, gi = base
            te_row = l2.time_embed.data_ptr() + t * D * _F32 if use_te else None
            prog.add(lib.OP_SCATTER, _scatter(bs.data_ptr(), out.data_ptr(), M, D, add_row=te_row))
            lasts = [plan.last_hist_f[h]] + ([plan.last_hist_b[h]] if bi else [])
            ent_sets = [x.snapshot.node_ids for x in lasts if x is not None]
            if ent_sets:
                ents = torch.from_numpy(np.unique(np.concatenate(ent_sets)).astype(np.int32)).to(dev)
                nc = int(ents.shape[0])
                gic = torch.empty(nc, 2 * G, dtype=torch.float32, device=dev)
                outc = torch.empty(nc, D, dtype=torch.float32, device=dev)
                ga = lib.GatherArgs()
                ga.n, ga.d, ga.table, ga.index, ga.out = nc, 2 * G, gi.data_ptr(), ents.data_ptr(), gic.data_ptr()
                prog.add(lib.OP_GATHER, ga)
                prev_c = [maps["prev"][j][h][ents.long()].contiguous() for j in range(len(dirs))]
                prog.keepalive += [ents, gic, outc] + prev_c
                for j, d in enumerate(dirs):
                    prog.add(lib.OP_GRU, rt._gru(l2, rnns[j][1], rnns[j][0], (0, nc), gi=gic, gi_ld=2 * G, gi_off=j * G,
                                                 state=S, prev=prev_c[j].data_ptr(), dt=ones, out=outc,
                                                 te=use_te and j == len(dirs) - 1, accumulate=j > 0, row_time_scalar=t,
                                                 layer_name="layer_2"))
                prog.add(lib.OP_SCATTER, _scatter(outc.data_ptr(), out.data_ptr(), nc, D, dst_index=ents.data_ptr()))
        else:
            if enc.rec_only_last_layer:
                prog.add(lib.OP_LAYER, iso_layer(l1, E, False, h_out=y1))
                first = y1
            else:
                first = ws.get("iso_first", M * D)[:M * D].view(M, D)
                prev1 = S if gru else res.bufs["state1"]
                rec_iso(l1, "layer_1", E, prev1, first, False, use_te)
            rec_iso(l2, "layer_2", first, S, out, relu2, use_te)

    # rows of the entities active at t come from the graph pass (DynamicRGCN.py:62-63)
    inst = plan.final.instances[i]
    n = inst.n
    src_base = res.out.data_ptr() + (inst.row0 - plan.final.row0) * D * _F32
    ids = torch.from_numpy(inst.snapshot.node_ids.astype(np.int32)).to(dev)
    prog.keepalive.append(ids)
    prog.add(lib.OP_SCATTER, _scatter(src_base, out.data_ptr(), n, D, dst_index=ids.data_ptr()))
    prog.run()
    return out
