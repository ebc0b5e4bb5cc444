"""Launch-program builder: turns (model parameters, WindowPlan) into a static list of kernel
launches (lib.Program) executed by ONE ``temp_run_program`` C call.

Per forward the reference issues, for every time step, a python loop over graphs, ``dgl.batch``,
host->device copies, ~40 torch/DGL kernels and a dense history re-zero.  Here the
whole window batch is:

    [H2D of the packed plan]  ->  recurrence-free work for ALL snapshot instances at once
    (layer 1, and with --rec-only-last-layer also the layer-2 aggregation + self loop + GRU input
    gates)  ->  one small GRU launch per time step (only h0 . W_hh^T and the gates are serial).

Families (reference file:line of what each program restates):
    SRGCN                models/RGCN.py:154-159
    GRRGCN / RRGCN       models/RRGCN.py:192-204 driven by models/DynamicRGCN.py:156-174, 132-144
    BiGRRGCN / BiRRGCN   models/BiRRGCN.py:210-240 driven by models/BiDynamicRGCN.py:77-121, 151-163
    SARGCN / BiSARGCN    models/SARGCN.py:103-117 driven by models/SelfAttentionRGCN.py:86-120 and
                         models/BiSelfAttentionRGCN.py:25-46, 71-87
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional

import numpy as np
import torch

from . import lib
from .planner import WindowPlan

_F32 = 4


class Workspace(object):
    """Grow-only device buffers keyed by name (no allocation on the steady-state path)."""

    def __init__(self, device):
        self.device = device
        self._bufs: Dict[str, torch.Tensor] = {}
        self._pinned: Dict[str, torch.Tensor] = {}

    def get(self, name: str, numel: int, dtype=torch.float32) -> torch.Tensor:
        t = self._bufs.get(name)
        if t is None or t.numel() < numel or t.dtype != dtype:
            cap = int(numel * 1.25) + 64
            t = torch.empty(cap, dtype=dtype, device=self.device)
            self._bufs[name] = t
        return t

    def pinned(self, name: str, numel: int, dtype=torch.uint8) -> torch.Tensor:
        t = self._pinned.get(name)
        if t is None or t.numel() < numel or t.dtype != dtype:
            cap = int(numel * 1.25) + 64
            t = torch.empty(cap, dtype=dtype, pin_memory=True)
            self._pinned[name] = t
        return t


class PreparedWeights(object):
    """Kernel-layout views of the parameters: transposed GRU / attention weights ([in, out] row-major,
    concatenated along out where one chained GEMM feeds several consumers) and the (w, b) pair of the
    learnable decay.  Rebuilt only when a parameter's version counter changes."""

    def __init__(self, model):
        self.model = model
        self._cache = {}

    def _key(self, tensors):
        return tuple((t.data_ptr(), t._version) for t in tensors)

    def cat_t(self, name: str, weights: List[torch.Tensor]) -> torch.Tensor:
        """cat([w.T for w in weights], dim=1) -- [in, sum(out)] row-major."""
        key = self._key(weights)
        hit = self._cache.get(name)
        if hit is None or hit[0] != key:
            with torch.no_grad():
                val = torch.cat([w.detach().t() for w in weights], dim=1).contiguous()
            hit = (key, val)
            self._cache[name] = hit
        return hit[1]

    def packed(self, name: str, w_kn: torch.Tensor, src: List[torch.Tensor]) -> Optional[torch.Tensor]:
        """tcgen05 operand image of a [k, n] row-major matrix (temp_pack_weights; k = 128 for the 128-row tile kernel, any
        k % 4 == 0 up to 256 for the 64-row one), None when the shape has no tensor-core path (the launch then runs on
        the fp32 SIMT kernels)."""
        k, n = int(w_kn.shape[0]), int(w_kn.shape[1])
        nbytes = lib.load().temp_packed_weights_bytes(k, n)
        if nbytes <= 0:
            return None
        key = self._key(src)
        hit = self._cache.get(name)
        if hit is None or hit[0] != key:
            buf = torch.empty(nbytes, dtype=torch.uint8, device=w_kn.device)
            w = w_kn.detach().contiguous()
            lib.check(lib.load().temp_pack_weights(C.c_void_p(w.data_ptr()), k, n, C.c_void_p(buf.data_ptr()),
                                                   C.c_void_p(lib.current_stream())), "temp_pack_weights")
            hit = (key, buf, w)
            self._cache[name] = hit
        return hit[1]

    def packed_gru(self, name: str, whh_t: torch.Tensor, src: List[torch.Tensor]) -> Optional[torch.Tensor]:
        d = int(whh_t.shape[0])
        nbytes = lib.load().temp_packed_gru_bytes(d)
        if nbytes <= 0 or int(whh_t.shape[1]) != 3 * d:
            return None
        key = self._key(src)
        hit = self._cache.get(name)
        if hit is None or hit[0] != key:
            buf = torch.empty(nbytes, dtype=torch.uint8, device=whh_t.device)
            lib.check(lib.load().temp_pack_gru_weights(C.c_void_p(whh_t.data_ptr()), d, C.c_void_p(buf.data_ptr()),
                                                       C.c_void_p(lib.current_stream())), "temp_pack_gru_weights")
            hit = (key, buf)
            self._cache[name] = hit
        return hit[1]

    def cat(self, name: str, vecs: List[torch.Tensor]) -> torch.Tensor:
        key = self._key(vecs)
        hit = self._cache.get(name)
        if hit is None or hit[0] != key:
            with torch.no_grad():
                val = torch.cat([v.detach().reshape(-1) for v in vecs]).contiguous()
            hit = (key, val)
            self._cache[name] = hit
        return hit[1]


class EncodeResult(object):
    def __init__(self, plan: WindowPlan, out: torch.Tensor, state: Optional[torch.Tensor], program: lib.Program,
                 bufs: dict):
        self.plan, self.out, self.state, self.program, self.bufs = plan, out, state, program, bufs

    @property
    def per_graph(self):
        """Views into the runtime's workspace: valid until the next forward of the model."""
        return tuple(self.out.split(self.plan.final_sizes))

    def per_graph_copy(self):
        """Fresh tensors, like the reference returns (one device copy): what the reference-API calls hand out, so that a
        caller collecting embeddings over several batches (test.py, the analysis scripts) keeps them."""
        return tuple(self.out.clone().split(self.plan.final_sizes))


def _p(t, off_bytes=0):
    return None if t is None else C.c_void_p(t.data_ptr() + off_bytes)


class EncoderRuntime(object):
    """Owns the workspace + prepared weights of one model and builds / runs launch programs."""

    def __init__(self, model):
        self.model = model
        self.device = model.ent_embeds.device
        if self.device.type != "cuda":
            raise RuntimeError("temp_b200: the encoder runs on CUDA only (no CPU fallback); move the model to "
                               "a cuda device")
        lib.load()
        self.ws = Workspace(self.device)
        self.prep = PreparedWeights(model)
        self._chain_packed = {}
        self.fuse_scan = True      # consecutive GRU steps -> one persistent scan launch (chain-partitioned on the tcgen05 path)
        self._agg_rows = 1
        self._live = []
        self._staged = {}
        # tcgen05 path where the shapes allow it (layers: d % 4 == 0, d <= 256; tensor-memory scan: d == 128); False: fp32 SIMT kernels only.  The 3xTF32 tensor-core GEMMs
        # are ~5x noisier than fp32 FFMA arithmetic (7.6e-7 against 1.8e-7 relative to an fp64 evaluation on the bench
        # shape).  That is irrelevant for the torch GRU / linear cells, but the --type1 cell is torch.randn-initialised
        # (models/GRU_cell.py:12-15): pre-activations of magnitude ~10 make its recurrence ill-conditioned (fp32 itself is
        # only good to ~1.5e-4 after 8 steps), so that flag stays on the exact-fp32 kernels.
        self.use_tc = not bool(getattr(model.args, "type1", False))
        # post-ensemble / impute shells: the layer-2 launches also store their output BEFORE the cells (+ time embedding),
        # the "local" stream of models/RRGCN.py:227-233, into bufs["local"]
        self.want_local = False

    # ---- plan upload -----------------------------------------------------------------------------
    def stage_plan(self, plan: WindowPlan, program: lib.Program, tag: str = "plan"):
        """Packs the plan into pinned host memory and appends the single H2D copy to ``program``.
        Returns name -> device pointer (int) of every plan array."""
        self._agg_rows = max(int(plan.R), 1)
        self._plan = plan
        self._live = program.keepalive
        lay, total = plan.blob_layout()
        pending = self._staged.pop(tag, None)
        if pending is not None:
            pending.synchronize()           # the last program staged under this tag still reads the pinned buffer
        host = self.ws.pinned(tag + "_host", total)
        plan.to_blob(host.numpy())
        dev = self.ws.get(tag + "_dev", total, torch.uint8)
        program.add(lib.OP_H2D, lib.CopyArgs(dev.data_ptr(), host.data_ptr(), total))
        program.keepalive += [host, dev]
        program.h2d_bytes = getattr(program, "h2d_bytes", 0) + total
        return {name: dev.data_ptr() + off for name, (off, _) in lay.items()}

    def mark_run(self, tag: str = "plan") -> None:
        """Call after launching a program staged under ``tag`` that will NOT be re-run once the tag is staged again
        (``model.encode`` per call, the per-step encoder calls): the next ``stage_plan(tag)`` waits for this point
        before it overwrites the pinned staging buffer the program's H2D copy reads."""
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._staged[tag] = ev

    # ---- op constructors ---------------------------------------------------------------------------
    def _layer(self, layer, rows, dptr, *, x, x_is_embed, terms, act, h_out=None, chain=None, te_out=False,
               te_chain=False, graph=True, residual=False, row_time_scalar=None):
        m = self.model
        a = lib.RgcnLayerArgs()
        a.row0, a.row1, a.d = int(rows[0]), int(rows[1]), m.embed_size
        if graph:
            a.row_ptr = dptr["row_ptr"]
            a.e_src = dptr["e_src_ent"] if x_is_embed else dptr["e_src"]
            a.e_rel = dptr["e_rel"]
            a.norm = dptr["norm"]
            a.x = x.data_ptr()
            a.weight = layer.weight.data_ptr()
            a.n_bases, a.si, a.so = layer.num_bases, layer.submat_in, layer.submat_out
            if self.use_tc and lib.load().temp_packed_weights_bytes(m.embed_size, m.embed_size) > 0:
                agg = self.ws.get("agg", self._agg_rows * m.embed_size)
                self._live.append(agg)               # programs hold raw pointers: keep the buffer they were built on
                a.agg_scratch = agg.data_ptr()
                if "agg_rows" in dptr:          # work lists of the aggregation launch, cut to [row0, row1)
                    a.agg_lists = 1
                    for name in ("agg_rows", "agg_heavy"):
                        ids = getattr(self._plan, name + "_ids")
                        lo, hi = int(np.searchsorted(ids, a.row0)), int(np.searchsorted(ids, a.row1))
                        setattr(a, name, dptr[name] + 12 * lo)
                        setattr(a, "n_" + name, hi - lo)
        a.residual = int(residual)
        a.n_terms = len(terms)
        for i, t in enumerate(terms):
            a.terms[i] = t
        if layer.bias:
            a.h_bias = layer.h_bias.data_ptr()
        a.activation = lib.ACT_RELU if act else lib.ACT_NONE
        if te_out or te_chain:
            a.time_embed = layer.time_embed.data_ptr()
            if row_time_scalar is None:
                a.row_time = dptr["row_time"]
            else:
                a.row_time_scalar = int(row_time_scalar)
        a.te_out, a.te_chain = int(te_out), int(te_chain)
        if h_out is not None:
            a.h_out = h_out.data_ptr()
        if chain is not None:
            w, b, out, ld = chain[:4]
            pk = self._chain_packed.get(w.data_ptr())
            if pk is not None:
                a.chain_w_packed = pk.data_ptr()
            a.chain_w = w.data_ptr()
            a.chain_b = None if b is None else b.data_ptr()
            a.chain_out = out.data_ptr()
            a.chain_n, a.chain_ld = int(w.shape[1]), int(ld)
        a.inv_temperature = float(m.args.inv_temperature)
        return a

    def _term(self, a, w, index=None, dt=None, decay_wb=None):
        t = lib.DenseTerm()
        t.a, t.w = a.data_ptr(), w.data_ptr()
        if self.use_tc and dt is None and int(w.shape[0]) == int(w.shape[1]):
            pk = self.prep.packed("term.%d" % w.data_ptr(), w, [w])
            if pk is not None:
                t.w_packed = pk.data_ptr()
        t.a_index = index
        t.a_dt = dt
        t.decay_wb = None if decay_wb is None else decay_wb.data_ptr()
        return t

    def _decay_wb(self, layer, name):
        if not layer.learnable_lambda:
            return None
        return self.prep.cat(name + ".decay_wb", [layer.exponential_decay.weight, layer.exponential_decay.bias])

    def _gru(self, layer, rnn, rnn_name, rows, *, gi, gi_ld, gi_off, state, prev, dt, out, te, accumulate=False,
             dptr=None, row_time_scalar=None, layer_name="", part_col=0):
        m = self.model
        type1 = bool(getattr(m.args, "type1", False))
        a = lib.GruArgs()
        a.row0, a.row1, a.d = int(rows[0]), int(rows[1]), m.embed_size
        a.gi, a.gi_ld, a.gi_off = gi.data_ptr(), int(gi_ld), int(gi_off)
        a.state = None if state is None else state.data_ptr()
        a.prev_row, a.dt = prev, dt
        wb = self._decay_wb(layer, layer_name)
        a.decay_wb = None if wb is None else wb.data_ptr()
        a.inv_temperature = float(m.args.inv_temperature)
        whh = rnn.weight_hh if type1 else rnn.weight_hh_l0
        bhh = rnn.bias_hh if type1 else rnn.bias_hh_l0
        whh_t = self.prep.cat_t(layer_name + "." + rnn_name + ".whh_t", [whh])
        a.whh_t = whh_t.data_ptr()
        if self.use_tc:
            pk = self.prep.packed_gru(layer_name + "." + rnn_name + ".whh_packed", whh_t, [whh])
            if pk is not None:
                a.whh_packed = pk.data_ptr()
        a.b_hh = bhh.data_ptr()
        a.cell_type = lib.CELL_TYPE1 if type1 else lib.CELL_TORCH_GRU
        if te:
            a.time_embed = layer.time_embed.data_ptr()
            if row_time_scalar is None:
                a.row_time = dptr["row_time"]
            else:
                a.row_time_scalar = int(row_time_scalar)
        a.accumulate = int(accumulate)
        a.out = out.data_ptr()
        a.out_index_is_row = 1
        a.part_col = int(part_col)
        return a

    def _wih(self, layer_name, rnns):
        """chained-GEMM operands for the input half of one or two GRUs: ([D, sum 3D|D], bias)."""
        type1 = bool(getattr(self.model.args, "type1", False))
        ws = [(r.weight_ih if type1 else r.weight_ih_l0) for _, r in rnns]
        bs = [(r.bias_ih if type1 else r.bias_ih_l0) for _, r in rnns]
        tag = layer_name + "." + "+".join(n for n, _ in rnns)
        w = self.prep.cat_t(tag + ".wih_t", ws)
        self._chain_packed[w.data_ptr()] = self.prep.packed(tag + ".wih_packed", w, ws) if self.use_tc else None
        return w, self.prep.cat(tag + ".b_ih", bs)

    # ---- programs ----------------------------------------------------------------------------------
    def build(self, plan: WindowPlan, with_h2d: bool = True, tag: str = "plan") -> EncodeResult:
        m = self.model
        fam = m.family
        prog = lib.Program()
        dptr = self.stage_plan(plan, prog, tag)
        if not with_h2d:
            prog.ops = [o for o in prog.ops if o.kind != lib.OP_H2D]
        if fam == "static":
            return self._build_static(plan, prog, dptr)
        if fam == "attention":
            return self._build_attention(plan, prog, dptr)
        return self._build_recurrent(plan, prog, dptr)

    SCAN_BARRIER_WORDS = 2 + 16 * 1024      # two grid-barrier words + per-(step, 64-row tile) completion counters

    def scan_barrier(self) -> int:
        """Zeroed words for the cooperative scans at D != 128 (self-cleaning, see TempGruScanArgs.barrier / barrier_words):
        the grid barrier, and the per-tile completion counters of gru_scan_tcw_kernel."""
        if getattr(self, "_barrier", None) is None:
            self._barrier = torch.zeros(self.SCAN_BARRIER_WORDS, dtype=torch.int32, device=self.device)
        return self._barrier.data_ptr()

    def _build_static(self, plan, prog, dptr):
        m, D, R = self.model, self.model.embed_size, plan.R
        enc = m.ent_encoder
        h1 = self.ws.get("h1", R * D).view(-1)[:R * D].view(R, D)
        out = self.ws.get("out", R * D).view(-1)[:R * D].view(R, D)
        rows = (0, R)
        prog.add(lib.OP_LAYER, self._layer(enc.layer_1, rows, dptr, x=m.ent_embeds, x_is_embed=True, act=False,
                                           terms=[self._term(m.ent_embeds, enc.layer_1.loop_weight, index=dptr["ent_id"])],
                                           h_out=h1))
        prog.add(lib.OP_LAYER, self._layer(enc.layer_2, rows, dptr, x=h1, x_is_embed=False, act=True,
                                           terms=[self._term(h1, enc.layer_2.loop_weight)], h_out=out,
                                           te_out=enc.use_time_embedding))
        return EncodeResult(plan, out, None, prog, {"h1": h1})

    def build_sharded(self, plan: WindowPlan, shard, rank: int) -> EncodeResult:
        """The rank's share of a snapshot-sharded forward (temp_b200/sharding.py): ``res.programs`` = [both RGCN
        layers for the rank's block of snapshot instances, the GRU scan over the rank's chain partitions]; the caller
        runs the two exchanges in between / after."""
        m = self.model
        if not (m.family == "recurrent" and m.ent_encoder.rec_only_last_layer and m.args.module in ("GRRGCN", "BiGRRGCN")
                and self.use_tc and m.embed_size == 128 and self.fuse_scan):
            raise RuntimeError("temp_b200: snapshot sharding covers the --rec-only-last-layer GRU families at D = 128; "
                               "other configurations shard over target timestamps only (SURVEY.md section 8e)")
        plan.scan_parts = shard.parts                       # same partitions, rank-major order
        prog = lib.Program()
        dptr = self.stage_plan(plan, prog)
        res = self._build_recurrent(plan, prog, dptr, shard=shard, rank=rank)
        ops = res.program.ops
        first_scan = next(i for i, o in enumerate(ops) if o.kind in (lib.OP_GRU, lib.OP_GRU_SCAN))
        res.programs = []
        for part in (ops[:first_scan], ops[first_scan:]):
            pr = lib.Program()
            pr.ops = list(part)
            pr.keepalive = res.program.keepalive
            res.programs.append(pr)
        return res

    def _build_recurrent(self, plan, prog, dptr, shard=None, rank=0):
        m, D, R = self.model, self.model.embed_size, plan.R

        def mine(rows):                                     # a launch's row range cut to this rank's block
            if shard is None:
                return rows
            lo, hi = shard.rows_of(rank)
            return (max(rows[0], lo), max(min(rows[1], hi), max(rows[0], lo)))

        enc = m.ent_encoder
        l1, l2 = enc.layer_1, enc.layer_2
        gru = m.args.module in ("GRRGCN", "BiGRRGCN")
        bi = plan.bidirectional
        use_te = enc.use_time_embedding
        relu2 = bi                                         # BiRRGCN.py:202-203 vs RRGCN.py:186-187
        type1 = bool(getattr(m.args, "type1", False))
        G = D if type1 else 3 * D                          # gi width of one cell
        GL = G * (2 if bi else 1)                          # row pitch of gi: one cell, two for the Bi centre step
        final = plan.final
        h1 = self.ws.get("h1", R * D)[:R * D].view(R, D)
        S = self.ws.get("state", R * D)[:R * D].view(R, D)
        S1 = None
        bufs = {"h1": h1}
        local_kw = {}
        if self.want_local:
            if not gru:
                raise NotImplementedError("temp_b200: the local (post-ensemble) stream exists for the GRU flavours only")
            bufs["local"] = self.ws.get("local", R * D)[:R * D].view(R, D)
            self._live.append(self.ws.get("local", R * D))
            local_kw = dict(h_out=bufs["local"], te_out=use_te)

        def rnn_of(layer, direction):
            if not bi:
                return ("rnn", layer.rnn) if gru else ("time_weight", layer.time_weight)
            if gru:
                return ("forward_rnn", layer.forward_rnn) if direction == "f" else ("backward_rnn", layer.backward_rnn)
            return (("time_weight_forward", layer.time_weight_forward) if direction == "f"
                    else ("time_weight_backward", layer.time_weight_backward))

        def dirs_of(seg):
            return ["f", "b"] if (seg.kind == "final" and bi) else [seg.kind[-1] if seg.kind != "final" else "f"]

        def prev_ptrs(direction, seg):
            which = "b" if (direction == "b" and seg.kind == "final") else "a"
            return dptr["prev_" + which], dptr["dt_" + which]

        # ---- recurrent layer over one row range (segment): GRU flavour or linear flavour ------------
        def rec_layer(layer, lname, rows, seg, x_in, x_is_embed, index, state_prev, out, relu, te):
            dirs = dirs_of(seg)
            if gru:
                rnns = [rnn_of(layer, d) for d in dirs]
                w, b = self._wih(lname, rnns)
                gi = self.ws.get("gi_" + lname, R * GL)[:R * GL].view(R, GL)
                self._live.append(gi)                    # a kept program must keep the buffers it addresses (grow-only workspace)
                prog.add(lib.OP_LAYER, self._layer(layer, rows, dptr, x=x_in, x_is_embed=x_is_embed, act=relu,
                                                   terms=[self._term(x_in, layer.loop_weight, index=index)],
                                                   chain=(w, b, gi, GL), **(local_kw if layer is l2 else {})))
                for j, d in enumerate(dirs):
                    pv, dt = prev_ptrs(d, seg)
                    prog.add(lib.OP_GRU, self._gru(layer, rnns[j][1], rnns[j][0], rows, gi=gi, gi_ld=GL, gi_off=j * G,
                                                   state=state_prev, prev=pv, dt=dt, out=out,
                                                   te=te and j == len(dirs) - 1, accumulate=j > 0, dptr=dptr,
                                                   layer_name=lname))
            else:
                terms = [self._term(x_in, layer.loop_weight, index=index)]
                for d in dirs:
                    pv, dt = prev_ptrs(d, seg)
                    terms.append(self._term(state_prev, rnn_of(layer, d)[1], index=pv, dt=dt))
                prog.add(lib.OP_LAYER, self._layer(layer, rows, dptr, x=x_in, x_is_embed=x_is_embed, act=relu, terms=terms,
                                                   h_out=out, te_out=te))

        if enc.rec_only_last_layer:
            # layer 1 for every snapshot instance at once
            prog.add(lib.OP_LAYER, self._layer(l1, mine((0, R)), dptr, x=m.ent_embeds, x_is_embed=True, act=False,
                                               terms=[self._term(m.ent_embeds, l1.loop_weight, index=dptr["ent_id"])],
                                               h_out=h1))
            if gru:
                # layer-2 aggregation + self loop + GRU input gates: also recurrence free.  Row groups
                # that share the same chained weights are launched together.
                gi = self.ws.get("gi_l2", R * GL)[:R * GL].view(R, GL)
                bufs["gi"] = gi
                groups = {}
                for seg in plan.segments:
                    key = tuple(dirs_of(seg))
                    lo, hi = groups.get(key, (seg.row0, seg.row1))
                    groups[key] = (min(lo, seg.row0), max(hi, seg.row1))
                for dirs, rows in groups.items():
                    rnns = [rnn_of(l2, d) for d in dirs]
                    w, b = self._wih("layer_2", rnns)
                    prog.add(lib.OP_LAYER, self._layer(l2, mine(rows), dptr, x=h1, x_is_embed=False, act=relu2,
                                                       terms=[self._term(h1, l2.loop_weight)], chain=(w, b, gi, GL), **local_kw))
                protos = {}          # one fully built GruArgs per (cell, variant); the steps copy it and patch the rows
                # Bi: the forward chain (forward history, then the centre step with the forward cell) before the backward
                # chain (backward history, centre step with the backward cell, ACCUMULATING into the forward result): two
                # runs of one cell each, i.e. two launches of the single-cell scan kernel instead of one two-cell scan
                steps = [(g, seg, j, d) for g, seg in enumerate(plan.segments) for j, d in enumerate(dirs_of(seg))]
                if bi:
                    steps = [s_ for s_ in steps if s_[3] == "f"] + [s_ for s_ in steps if s_[3] == "b"]
                for g, seg, j, d in steps:
                    dirs = dirs_of(seg)
                    if True:
                        pv, dt = prev_ptrs(d, seg)
                        key = (d, j, len(dirs), pv, dt)
                        proto = protos.get(key)
                        if proto is None:
                            name, rnn = rnn_of(l2, d)
                            proto = self._gru(l2, rnn, name, (0, 0), gi=gi, gi_ld=GL, gi_off=j * G, state=S, prev=pv, dt=dt,
                                              out=S, te=use_te and j == len(dirs) - 1, accumulate=j > 0, dptr=dptr,
                                              layer_name="layer_2")
                            protos[key] = proto
                        a = lib.GruArgs.from_buffer_copy(proto)
                        a.row0, a.row1, a.part_col = seg.row0, seg.row1, g
                        prog.add(lib.OP_GRU, a)
            else:
                for seg in plan.segments:
                    rec_layer(l2, "layer_2", (seg.row0, seg.row1), seg, h1, False, None, S, S, relu2, use_te)
        else:
            # both layers recurrent: strictly serial in the step index.  For the GRU flavours the
            # reference's graph aliasing feeds layer 1 with the previous LAYER-2 state (Appendix B-2);
            # the linear flavours keep separate first / second states.
            S1 = self.ws.get("state1", R * D)[:R * D].view(R, D)
            bufs["state1"] = S1
            prev1 = S if gru else S1
            for seg in plan.segments:
                rows = (seg.row0, seg.row1)
                rec_layer(l1, "layer_1", rows, seg, m.ent_embeds, True, dptr["ent_id"], prev1, S1, False, use_te)
                rec_layer(l2, "layer_2", rows, seg, S1, False, None, S, S, relu2, use_te)
        out = S[final.row0:final.row1]
        bufs["state"] = S
        if self.fuse_scan:
            parts = plan.scan_parts if (self.use_tc and enc.rec_only_last_layer and gru) else None
            if parts is not None and parts.shape[0] > 0:
                p_lo, p_hi = (0, int(parts.shape[0])) if shard is None else shard.parts_of(rank)
                stride = int(parts.shape[1])
                prog.fuse_gru_scans(self.scan_barrier(), dptr["scan_parts"] + 8 * stride * p_lo, p_hi - p_lo, stride,
                                    int(getattr(plan, "scan_tile", 0)), split_cells=(D == 128), barrier_words=self.SCAN_BARRIER_WORDS)   # (the SIMT scan takes both cells in one launch)
            else:
                prog.fuse_gru_scans(self.scan_barrier(), barrier_words=self.SCAN_BARRIER_WORDS)
        return EncodeResult(plan, out, S, prog, bufs)

    def _build_attention(self, plan, prog, dptr):
        m, D, R = self.model, self.model.embed_size, plan.R
        enc = m.ent_encoder
        l1, l2 = enc.layer_1, enc.layer_2
        final = plan.final
        Rh = final.row0
        nf = final.row1 - final.row0
        h1 = self.ws.get("h1", R * D)[:R * D].view(R, D)
        out = self.ws.get("out", max(nf, 1) * D)[:nf * D].view(nf, D)
        out_base = out.data_ptr() - final.row0 * D * _F32          # kernels address rows absolutely
        tau = torch.tensor(m.time_diff(plan), dtype=torch.float32, device=self.device)
        prog.keepalive.append(tau)

        def chains(layer, lname):
            kvw = [layer.k_linear.weight, layer.v_linear.weight]
            qkvw = [layer.q_linear.weight] + kvw
            kv = self.prep.cat_t(lname + ".kv_t", kvw)
            qkv = self.prep.cat_t(lname + ".qkv_t", qkvw)
            if self.use_tc:
                self._chain_packed[kv.data_ptr()] = self.prep.packed(lname + ".kv_packed", kv, kvw)
                self._chain_packed[qkv.data_ptr()] = self.prep.packed(lname + ".qkv_packed", qkv, qkvw)
            kvb = self.ws.get("kv_" + lname, max(Rh, 1) * 2 * D)[:Rh * 2 * D].view(Rh, 2 * D)
            qkvb = self.ws.get("qkv_" + lname, max(nf, 1) * 3 * D)[:nf * 3 * D].view(nf, 3 * D)
            prog.keepalive += [kvb, qkvb]                # a kept program must keep the buffers it addresses
            return kv, qkv, kvb, qkvb

        def attend(layer, lname, kvb, qkvb, combine):
            a = lib.AttnArgs()
            a.row0, a.row1, a.d, a.heads = final.row0, final.row1, D, layer.h
            a.qkv = qkvb.data_ptr() - final.row0 * 3 * D * _F32
            a.kv_hist = kvb.data_ptr() if Rh > 0 else None
            a.slot_row = dptr["slot_row"] - final.row0 * plan.n_slots * 4 if plan.n_slots > 0 else None
            a.n_slots = plan.n_slots if Rh > 0 else 0
            a.tau = tau.data_ptr()
            wb = self._decay_wb(layer, lname)
            a.decay_wb = None if wb is None else wb.data_ptr()
            a.combine_max = int(combine)
            a.out = out_base
            prog.add(lib.OP_ATTN, a)

        class _Abs(object):   # tensor stand-in whose data_ptr is shifted so that row r maps to r - row0
            def __init__(self, t, shift):
                self._p = t.data_ptr() - shift
            def data_ptr(self):
                return self._p

        bufs = {"h1": h1}
        emb_term = [self._term(m.ent_embeds, l1.loop_weight, index=dptr["ent_id"])]
        if enc.rec_only_last_layer:
            prog.add(lib.OP_LAYER, self._layer(l1, (0, R), dptr, x=m.ent_embeds, x_is_embed=True, act=False,
                                               terms=emb_term, h_out=h1))
        else:
            kv1, qkv1, kvb1, qkvb1 = chains(l1, "layer_1")
            if Rh > 0:
                prog.add(lib.OP_LAYER, self._layer(l1, (0, Rh), dptr, x=m.ent_embeds, x_is_embed=True, act=False,
                                                   terms=emb_term, h_out=h1, te_chain=True,
                                                   chain=(kv1, None, kvb1, 2 * D)))
            prog.add(lib.OP_LAYER, self._layer(l1, (Rh, R), dptr, x=m.ent_embeds, x_is_embed=True, act=False,
                                               terms=emb_term, h_out=h1, te_chain=True,
                                               chain=(qkv1, None, _Abs(qkvb1, Rh * 3 * D * _F32), 3 * D)))
            attend(l1, "layer_1", kvb1, qkvb1, combine=False)
            bufs["kv1"] = kvb1
        kv2, qkv2, kvb2, qkvb2 = chains(l2, "layer_2")
        t2 = [self._term(h1, l2.loop_weight)]
        if Rh > 0:
            prog.add(lib.OP_LAYER, self._layer(l2, (0, Rh), dptr, x=h1, x_is_embed=False, act=True, terms=t2,
                                               te_chain=True, chain=(kv2, None, kvb2, 2 * D)))
        prog.add(lib.OP_LAYER, self._layer(l2, (Rh, R), dptr, x=h1, x_is_embed=False, act=True, terms=t2,
                                           te_chain=True, chain=(qkv2, None, _Abs(qkvb2, Rh * 3 * D * _F32), 3 * D)))
        attend(l2, "layer_2", kvb2, qkvb2, combine=not enc.rec_only_last_layer)   # JK max, SARGCN.py:117
        bufs.update(kv2=kvb2, qkv2=qkvb2)
        return EncodeResult(plan, out, None, prog, bufs)
