"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the TeMP RGCN + GRU/BiGRU/attention forward.

This file restates, in plain torch-CPU fp32 ops and without DGL / pytorch-lightning, the
algorithm of the reference path named by BASELINE.json (JiapengWu/TeMP @ 53e90c1).  It is the
checker for the CUDA path (tests/, __graft_entry__.smoke) and the timed ``cpu_baseline`` arm of
bench.py.  Nothing under temp_b200/ may import it.

PARITY PIN: the reference ships no tests or golden vectors, so the pin is
"outputs of the reference itself run here": tests/golden/make_golden.py imports the UNMODIFIED
reference python files from /root/reference on top of oracle/_stubs (a restatement of the DGL
0.4.1 / pytorch-lightning 0.5.2 API slices the reference calls) and stores its outputs under
tests/golden/*.npz; tests/test_oracle_golden.py checks every function here against them.  The
aggregation arithmetic itself lives in DGL 0.4.1 (``update_all`` + ``fn.sum``), absent from
/root/reference; it is restated as a zero-initialised ``index_add_`` over edges in edge-id order.

The op sequence deliberately keeps the reference's orchestration costs (per-time-step batching,
per-graph python loops, dense (B, 2, M, D) history re-zeroed every step), because this file is
also the honest CPU baseline.

Layout conventions
  graph  : SnapGraph(ids[N] int64 global entity ids, src/dst/rel[E] int64 local ids, norm[N] f32)
  params : dict name -> tensor using the reference's state_dict key names
  cfg    : OracleConfig
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------------
# configuration
# --------------------------------------------------------------------------------------------
@dataclass
class OracleConfig:
    module: str = "GRRGCN"           # GRRGCN | RRGCN | BiGRRGCN | BiRRGCN | SARGCN | BiSARGCN | SRGCN
    num_ents: int = 0
    num_rels: int = 0                # layers allocate 2*num_rels relation rows (models/RGCN.py:149-152)
    num_times: int = 0               # rows of every per-layer time_embed (models/RGCN.py:15)
    embed_size: int = 128            # == hidden_size
    n_bases: int = 128
    seq_len: int = 8                 # train_seq_len == test_seq_len (models/DynamicRGCN.py:17-18)
    rec_only_last_layer: bool = True
    use_time_embedding: bool = True
    type1: bool = False
    learnable_lambda: bool = False
    inv_temperature: float = 0.1
    heads: int = 8                   # models/SARGCN.py:20
    use_embed_for_non_active: bool = False   # get_all_embeds_Gt keeps ent_embeds for entities without an edge at t

    @property
    def bidirectional(self) -> bool:
        return self.module.startswith("Bi")

    @property
    def attention(self) -> bool:
        return self.module in ("SARGCN", "BiSARGCN")

    @property
    def gru(self) -> bool:
        return self.module in ("GRRGCN", "BiGRRGCN")

    @property
    def layer_bias(self) -> bool:
        # RRGCN/BiRRGCN layers: bias=False (models/RRGCN.py:180-187, BiRRGCN.py:196-203);
        # RGCN (SRGCN) and SARGCN layers keep the RGCNLayer default bias=True
        # (models/RGCN.py:149-152, models/SARGCN.py:94-101).
        return self.module in ("SRGCN", "SARGCN", "BiSARGCN")

    @property
    def layer2_relu(self) -> bool:
        # relu on layer 2 for RGCN, SARGCN and BiRRGCN (models/RGCN.py:151, SARGCN.py:100,
        # BiRRGCN.py:202-203); None for the uni-directional RRGCN (models/RRGCN.py:186-187).
        return self.module in ("SRGCN", "SARGCN", "BiSARGCN", "BiGRRGCN", "BiRRGCN")


@dataclass
class SnapGraph:
    ids: np.ndarray                  # [N] global entity id of local node i (sorted unique)
    src: np.ndarray                  # [E] local
    dst: np.ndarray                  # [E] local
    rel: np.ndarray                  # [E] in [0, num_rels)
    norm: np.ndarray                 # [N] float32 1/in_degree, 0 for in_degree 0
    time: int = -1

    @property
    def num_nodes(self) -> int:
        return int(self.ids.shape[0])

    @property
    def num_edges(self) -> int:
        return int(self.src.shape[0])


# --------------------------------------------------------------------------------------------
# data: quadruple files -> per-timestamp graphs   (utils/dataset.py:151-232, 235-252, 268-290)
# --------------------------------------------------------------------------------------------
def in_degree_norm(dst: np.ndarray, n: int) -> np.ndarray:
    """utils/utils.py:74-79 (comp_deg_norm): 1/in_deg as float32, inf -> 0."""
    deg = np.bincount(dst, minlength=n).astype(np.float32)
    with np.errstate(divide="ignore"):
        norm = (1.0 / deg).astype(np.float32)
    norm[np.isinf(norm)] = 0
    return norm


def read_stat(dataset_path: str) -> Tuple[int, int]:
    """utils/dataset.py:56-60."""
    with open(os.path.join(dataset_path, "stat.txt")) as fr:
        parts = fr.readline().split()
    return int(parts[0]), int(parts[1])


def read_quadruples_by_time(dataset_path: str):
    """utils/dataset.py:235-252: time -> {train, valid, test} lists of (h, r, t) in file order."""
    per_time: Dict[int, Dict[str, list]] = {}
    for fname, mode in (("train.txt", "train"), ("valid.txt", "valid"), ("test.txt", "test")):
        with open(os.path.join(dataset_path, fname)) as fr:
            for line in fr:
                p = line.split()
                if not p:
                    continue
                h, r, t, tim = int(p[0]), int(p[1]), int(p[2]), int(p[3])
                per_time.setdefault(tim, {"train": [], "valid": [], "test": []})[mode].append((h, r, t))
    return dict(sorted(per_time.items()))


def graphs_at_time(triples: Dict[str, list], tim: int) -> Tuple[SnapGraph, SnapGraph, SnapGraph]:
    """utils/dataset.py:151-232 with add_reverse=False (line 186): the node set is the sorted
    unique of all train+valid+test subjects and objects at this time, shared by the three graphs;
    edges keep file order; relation ids stay in [0, num_rels)."""
    parts = [np.asarray(triples[m], dtype=np.int64).reshape(-1, 3) for m in ("train", "valid", "test")]
    total = np.concatenate(parts, axis=0)
    uniq = np.unique(np.concatenate([total[:, 0], total[:, 2]]))
    out, lo = [], 0
    for part in parts:
        hi = lo + part.shape[0]
        src = np.searchsorted(uniq, total[lo:hi, 0])
        dst = np.searchsorted(uniq, total[lo:hi, 2])
        rel = total[lo:hi, 1].copy()
        out.append(SnapGraph(ids=uniq.copy(), src=src, dst=dst, rel=rel,
                             norm=in_degree_norm(dst, uniq.shape[0]), time=tim))
        lo = hi
    return out[0], out[1], out[2]


def build_graph_dicts(dataset_path: str):
    """utils/dataset.py:268-290 without the pickle cache: three dicts time -> SnapGraph."""
    per_time = read_quadruples_by_time(dataset_path)
    train, valid, test = {}, {}, {}
    for tim, tr in per_time.items():
        train[tim], valid[tim], test[tim] = graphs_at_time(tr, tim)
    return train, valid, test


def edge_subgraph(g: SnapGraph, edge_idx: np.ndarray) -> SnapGraph:
    """models/DynamicRGCN.py:82-89: edge_subgraph(preserve_nodes=True) + recomputed norms."""
    src, dst, rel = g.src[edge_idx], g.dst[edge_idx], g.rel[edge_idx]
    return SnapGraph(ids=g.ids, src=src, dst=dst, rel=rel, norm=in_degree_norm(dst, g.num_nodes), time=g.time)


# --------------------------------------------------------------------------------------------
# layers
# --------------------------------------------------------------------------------------------
def sample_edges(g: SnapGraph, rate: float) -> SnapGraph:
    """models/DynamicRGCN.py:82-89: np.random.choice(arange(E), size=int(rate * E), replace=False) -> edge_subgraph with
    the drawn order kept and norms recomputed."""
    idx = np.random.choice(np.arange(g.num_edges), size=int(rate * g.num_edges), replace=False)
    return edge_subgraph(g, idx)


def batch_graphs(graphs: Sequence[SnapGraph]):
    """dgl.batch as used at models/DynamicRGCN.py:92: block-diagonal union, node ids offset by the
    running node count, edges concatenated in list order."""
    srcs, dsts, rels, ids, norms, off = [], [], [], [], [], 0
    for g in graphs:
        srcs.append(g.src + off)
        dsts.append(g.dst + off)
        rels.append(g.rel)
        ids.append(g.ids)
        norms.append(g.norm)
        off += g.num_nodes
    as_t = lambda xs, dt: torch.from_numpy(np.concatenate(xs).astype(dt)) if xs else torch.zeros(0)
    return (as_t(srcs, np.int64), as_t(dsts, np.int64), as_t(rels, np.int64),
            as_t(ids, np.int64), as_t(norms, np.float32))


def rgcn_aggregate(weight: Tensor, n_bases: int, h: Tensor, src: Tensor, dst: Tensor, rel: Tensor,
                   node_norm: Tensor) -> Tensor:
    """models/RGCN.py:91-104.  msg_func: gather W[rel] viewed [nb*?, si, so] and bmm with h[src]
    viewed [-1, 1, si]; times the edge norm (= node norm of dst, utils/utils.py:23-28); fn.sum into
    dst rows (zero rows for in-degree 0); apply_func multiplies by the node norm again."""
    d_in = h.shape[1]
    si = d_in // n_bases
    so = weight.shape[1] // (n_bases * si)
    d_out = n_bases * so
    w = weight.index_select(0, rel).view(-1, si, so)
    node = h[src].view(-1, 1, si)
    msg = torch.bmm(node, w).view(-1, d_out)
    msg = msg * node_norm[dst].view(-1, 1)
    agg = torch.zeros(h.shape[0], d_out, dtype=h.dtype).index_add_(0, dst, msg)
    return agg * node_norm.view(-1, 1)


def rgcn_layer_graph(p: Dict[str, Tensor], pre: str, cfg: OracleConfig, h: Tensor, bg, relu: bool) -> Tensor:
    """models/RGCN.py:53-76 in eval mode (dropout is the identity)."""
    src, dst, rel, _, norm = bg
    loop = torch.mm(h, p[pre + "loop_weight"])
    out = rgcn_aggregate(p[pre + "weight"], cfg.n_bases, h, src, dst, rel, norm)
    if cfg.layer_bias:
        out = out + p[pre + "h_bias"]
    out = out + loop
    return torch.relu(out) if relu else out


def rgcn_layer_isolated(p: Dict[str, Tensor], pre: str, cfg: OracleConfig, x: Tensor, relu: bool) -> Tensor:
    """models/RGCN.py:78-89: note the residual ``x +`` that the graph path lacks."""
    x = x + torch.mm(x, p[pre + "loop_weight"])
    if cfg.layer_bias:
        x = x + p[pre + "h_bias"]
    return torch.relu(x) if relu else x


def time_rows(p: Dict[str, Tensor], pre: str, times: Sequence[int], sizes: Sequence[int]) -> Tensor:
    """models/RGCN.py:47-51: per-graph time embedding row broadcast to the graph's nodes."""
    te = p[pre + "time_embed"]
    return torch.cat([te[int(t)].unsqueeze(0).expand(n, te.shape[1]) for t, n in zip(times, sizes)], dim=0)


def decay_state(p: Dict[str, Tensor], pre: str, cfg: OracleConfig, prev: Tensor, dt: Tensor) -> Tensor:
    """models/RRGCN.py:79-83 / models/RGCN.py:106-107."""
    if cfg.learnable_lambda:
        lam = dt * p[pre + "exponential_decay.weight"].view(1, 1) + p[pre + "exponential_decay.bias"].view(1, 1)
        return prev * torch.exp(-torch.clamp(lam, min=0))
    return prev * torch.exp(-dt * cfg.inv_temperature)


def gru_step(p: Dict[str, Tensor], pre: str, cfg: OracleConfig, x: Tensor, h0: Tensor) -> Tensor:
    """torch.nn.GRU (1 layer, seq_len 1, gate order r,z,n) as used at
    models/RRGCN.py:84, or the hand-written ``type1`` cell of models/GRU_cell.py:18-31."""
    D = h0.shape[1]
    if cfg.type1:
        i_n = torch.mm(x, p[pre + "weight_ih"].t()) + p[pre + "bias_ih"]
        gh = torch.mm(h0, p[pre + "weight_hh"].t()) + p[pre + "bias_hh"]
        h_r, h_i, h_n = gh[:, :D], gh[:, D:2 * D], gh[:, 2 * D:]
        r, z = torch.sigmoid(h_r), torch.sigmoid(h_i)
        n = torch.tanh(i_n + r * h_n)
        return n + z * (h0 - n)
    gi = torch.mm(x, p[pre + "weight_ih_l0"].t()) + p[pre + "bias_ih_l0"]
    gh = torch.mm(h0, p[pre + "weight_hh_l0"].t()) + p[pre + "bias_hh_l0"]
    r = torch.sigmoid(gi[:, :D] + gh[:, :D])
    z = torch.sigmoid(gi[:, D:2 * D] + gh[:, D:2 * D])
    n = torch.tanh(gi[:, 2 * D:] + r * gh[:, 2 * D:])
    return (1 - z) * n + z * h0


def attention_mix(p: Dict[str, Tensor], pre: str, cfg: OracleConfig, cur: Tensor, prev: Tensor,
                  tau: Tensor, mask: Tensor) -> Tensor:
    """models/SARGCN.py:25-53.  cur [N,D], prev [N,Lp,D], mask [N,Lp+1], tau [Lp+1].
    Output channel order is [d_k major, head minor]."""
    N, D = cur.shape
    H, dk = cfg.heads, D // cfg.heads
    if cfg.learnable_lambda:
        lam = tau.view(-1, 1) * p[pre + "exponential_decay.weight"].view(1, 1) + p[pre + "exponential_decay.bias"].view(1, 1)
        dec = -torch.clamp(lam, min=0).view(-1)
    else:
        dec = 0
    allv = torch.cat([prev, cur.unsqueeze(1)], dim=1)
    q = torch.mm(cur, p[pre + "q_linear.weight"].t()).view(N, 1, H, dk).transpose(1, 2)
    k = torch.matmul(allv, p[pre + "k_linear.weight"].t()).view(N, -1, H, dk).transpose(1, 2)
    v = torch.matmul(allv, p[pre + "v_linear.weight"].t()).view(N, -1, H, dk).transpose(1, 2)
    sc = torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(dk)          # [N,H,1,S]
    sc = sc.view(N, H, -1) + mask.unsqueeze(1) + dec
    w = torch.softmax(sc, dim=-1)
    o = torch.matmul(w.unsqueeze(2), v).view(N, H, dk)                  # [N,H,dk]
    return o.transpose(1, 2).contiguous().view(N, D)


# --------------------------------------------------------------------------------------------
# window construction   (models/TKG_Module.py:232-250, models/BiDynamicRGCN.py:17-49)
# --------------------------------------------------------------------------------------------
def window_forward(t_list: Sequence[int], seq_len: int, times: List[int]):
    """Returns time_batched[k][i] (None padded at the window front); items sorted descending."""
    order = sorted([int(t) for t in t_list], reverse=True)
    rows = []
    for tim in order:
        length = times.index(tim) + 1
        seq = times[length - seq_len:length] if seq_len <= length else times[:length]
        rows.append([None] * (seq_len - len(seq)) + list(seq))
    return [list(x) for x in zip(*rows)]


def window_backward(t_list: Sequence[int], seq_len: int, times: List[int]):
    """models/BiDynamicRGCN.py:36-41: items sorted ASCENDING, window [t .. t+L-1] reversed so that
    t is last, None padded at the front."""
    order = sorted([int(t) for t in t_list])
    rows, T = [], len(times)
    for tim in order:
        pos = times.index(tim)
        seq = times[pos:pos + seq_len] if seq_len <= T - pos else times[pos:]
        seq = list(reversed(seq))
        rows.append([None] * (seq_len - len(seq)) + seq)
    return [list(x) for x in zip(*rows)]


# --------------------------------------------------------------------------------------------
# the model
# --------------------------------------------------------------------------------------------
class OracleModel:
    """Functional restatement of the reference model shells' deterministic (eval-mode) forward:
    DynamicRGCN / BiDynamicRGCN / SelfAttentionRGCN / BiSelfAttentionRGCN / StaticRGCN."""

    def __init__(self, cfg: OracleConfig, params: Dict[str, Tensor], graph_dict_train: Dict[int, SnapGraph]):
        self.cfg, self.p, self.gd = cfg, params, graph_dict_train
        self.times = list(graph_dict_train.keys())
        self.M, self.D = cfg.num_ents, cfg.embed_size

    # ---- encoders on a batched graph ---------------------------------------------------------
    def _embed(self, bg) -> Tensor:
        return self.p["ent_embeds"][bg[3]]                                # DynamicRGCN.py:93

    def enc_static(self, graphs, times) -> Tensor:
        """models/RGCN.py:154-159."""
        cfg, p = self.cfg, self.p
        bg = batch_graphs(graphs)
        sizes = [g.num_nodes for g in graphs]
        h1 = rgcn_layer_graph(p, "ent_encoder.layer_1.", cfg, self._embed(bg), bg, relu=False)
        h2 = rgcn_layer_graph(p, "ent_encoder.layer_2.", cfg, h1, bg, relu=True)
        if cfg.use_time_embedding:
            h2 = h2 + time_rows(p, "ent_encoder.layer_2.", times, sizes)
        return h2

    def enc_static_isolated(self, t: int) -> Tensor:
        """models/RGCN.py:161-164."""
        cfg, p = self.cfg, self.p
        x = rgcn_layer_isolated(p, "ent_encoder.layer_1.", cfg, p["ent_embeds"], relu=False)
        x = rgcn_layer_isolated(p, "ent_encoder.layer_2.", cfg, x, relu=True)
        return x + p["ent_encoder.layer_2.time_embed"][int(t)] if cfg.use_time_embedding else x

    def _rec_layer_graph(self, pre, h, bg, prevs, dts, rnn_names, relu):
        """One recurrent layer on graph rows.  GRU flavour: models/RRGCN.py:77-89,
        BiRRGCN.py:27-63; linear flavour: models/RRGCN.py:130-154, BiRRGCN.py:115-164.
        prevs/dts/rnn_names are parallel lists (one entry uni / one-direction, two for the Bi
        centre step)."""
        cfg, p = self.cfg, self.p
        if cfg.gru:
            x = rgcn_layer_graph(p, pre, cfg, h, bg, relu)
            out = 0
            for prev, dt, rn in zip(prevs, dts, rnn_names):
                out = out + gru_step(p, pre + rn + ".", cfg, x, decay_state(p, pre, cfg, prev, dt))
            return x, out
        src, dst, rel, _, norm = bg
        loop = torch.mm(h, p[pre + "loop_weight"])
        out = rgcn_aggregate(p[pre + "weight"], cfg.n_bases, h, src, dst, rel, norm)
        if len(prevs) == 1:      # RRGCN.py:142 / BiRRGCN.py:153: decay multiplies the projected state
            out = out + torch.mm(prevs[0], p[pre + rnn_names[0]]) * torch.exp(-dts[0] * cfg.inv_temperature)
        else:                    # BiRRGCN.py:124-129: decay applied before the projection
            for prev, dt, wn in zip(prevs, dts, rnn_names):
                out = out + torch.mm(prev * torch.exp(-dt * cfg.inv_temperature), p[pre + wn])
        out = out + loop
        return None, (torch.relu(out) if relu else out)

    def _rnn_names(self, direction: Optional[str]):
        cfg = self.cfg
        if cfg.module == "GRRGCN":
            return ["rnn"]
        if cfg.module == "RRGCN":
            return ["time_weight"]
        base = ("%s_rnn" if cfg.gru else "time_weight_%s")
        if direction is None:
            return [base % "forward", base % "backward"]
        return [base % direction]

    def enc_recurrent(self, graphs, times, prev1, prev2, dts, direction: Optional[str] = "forward"):
        """models/RRGCN.py:192-204 and models/BiRRGCN.py:210-240.  prev1/prev2/dts are lists with one
        entry per direction used.  Returns (first, second) with the graph-aliasing quirk of
        SURVEY Appendix B-2: identical tensors for the GRU flavours."""
        cfg, p = self.cfg, self.p
        bg = batch_graphs(graphs)
        sizes = [g.num_nodes for g in graphs]
        names = self._rnn_names(direction)
        relu2 = cfg.layer2_relu
        h0 = self._embed(bg)
        if cfg.rec_only_last_layer:
            first = rgcn_layer_graph(p, "ent_encoder.layer_1.", cfg, h0, bg, relu=False)
        else:
            _, first = self._rec_layer_graph("ent_encoder.layer_1.", h0, bg, prev1, dts, names, relu=False)
            if cfg.use_time_embedding:
                first = first + time_rows(p, "ent_encoder.layer_1.", times, sizes)
        _, second = self._rec_layer_graph("ent_encoder.layer_2.", first, bg, prev2, dts, names, relu=relu2)
        if cfg.use_time_embedding:
            second = second + time_rows(p, "ent_encoder.layer_2.", times, sizes)
        if cfg.gru:
            first = second                                                # Appendix B-2
        return first, second

    def enc_recurrent_isolated(self, t: int, prev1, prev2, dts):
        """models/RRGCN.py:206-217, models/BiRRGCN.py:242-257 (+ layer code RRGCN.py:91-104, 156-167,
        BiRRGCN.py:65-81, 166-185).  All-entity rows; prev*/dts are lists per direction."""
        cfg, p = self.cfg, self.p
        names = self._rnn_names(None if cfg.bidirectional else "forward")

        def rec_iso(pre, x, prevs, relu):
            if cfg.gru:
                x = rgcn_layer_isolated(p, pre, cfg, x, relu)
                out = 0
                for prev, dt, rn in zip(prevs, dts, names):
                    out = out + gru_step(p, pre + rn + ".", cfg, x, decay_state(p, pre, cfg, prev, dt))
                return out
            x = x + torch.mm(x, p[pre + "loop_weight"])
            if len(prevs) == 1:
                x = x + torch.mm(prevs[0], p[pre + names[0]]) * torch.exp(-dts[0] * cfg.inv_temperature)
            else:
                for prev, dt, wn in zip(prevs, dts, names):
                    x = x + torch.mm(prev * torch.exp(-dt * cfg.inv_temperature), p[pre + wn])
            return torch.relu(x) if relu else x

        x = p["ent_embeds"]
        if cfg.rec_only_last_layer:
            first = rgcn_layer_isolated(p, "ent_encoder.layer_1.", cfg, x, relu=False)
        else:
            first = rec_iso("ent_encoder.layer_1.", x, prev1, relu=False)
            if cfg.use_time_embedding:
                first = first + p["ent_encoder.layer_1.time_embed"][int(t)]
        second = rec_iso("ent_encoder.layer_2.", first, prev2, relu=cfg.layer2_relu)
        if cfg.use_time_embedding:
            second = second + p["ent_encoder.layer_2.time_embed"][int(t)]
        return second

    def enc_attention_history(self, graphs, times):
        """models/SARGCN.py:103-107: two plain layers; both outputs get their time embedding added,
        layer 2 consumes layer 1 WITHOUT it."""
        cfg, p = self.cfg, self.p
        bg = batch_graphs(graphs)
        sizes = [g.num_nodes for g in graphs]
        h1 = rgcn_layer_graph(p, "ent_encoder.layer_1.", cfg, self._embed(bg), bg, relu=False)
        h2 = rgcn_layer_graph(p, "ent_encoder.layer_2.", cfg, h1, bg, relu=True)
        return (h1 + time_rows(p, "ent_encoder.layer_1.", times, sizes),
                h2 + time_rows(p, "ent_encoder.layer_2.", times, sizes))

    def enc_attention_final(self, graphs, times, prev1, prev2, tau, mask):
        """models/SARGCN.py:38-47, 109-117."""
        cfg, p = self.cfg, self.p
        bg = batch_graphs(graphs)
        sizes = [g.num_nodes for g in graphs]
        h1 = rgcn_layer_graph(p, "ent_encoder.layer_1.", cfg, self._embed(bg), bg, relu=False)
        if not cfg.rec_only_last_layer:
            cur1 = h1 + time_rows(p, "ent_encoder.layer_1.", times, sizes)
            a1 = attention_mix(p, "ent_encoder.layer_1.", cfg, cur1, prev1, tau, mask)
        h2 = rgcn_layer_graph(p, "ent_encoder.layer_2.", cfg, h1, bg, relu=True)
        cur2 = h2 + time_rows(p, "ent_encoder.layer_2.", times, sizes)
        a2 = attention_mix(p, "ent_encoder.layer_2.", cfg, cur2, prev2, tau, mask)
        return a2 if cfg.rec_only_last_layer else torch.max(a1, a2)

    def enc_attention_isolated(self, t: int, prev1, prev2, tau, mask):
        """models/SARGCN.py:55-62, 119-125."""
        cfg, p = self.cfg, self.p
        x = p["ent_embeds"]
        if cfg.rec_only_last_layer:
            first = rgcn_layer_isolated(p, "ent_encoder.layer_1.", cfg, x, relu=False)
        else:
            c1 = rgcn_layer_isolated(p, "ent_encoder.layer_1.", cfg, x, relu=False)
            first = attention_mix(p, "ent_encoder.layer_1.", cfg, c1 + p["ent_encoder.layer_1.time_embed"][int(t)],
                                  prev1, tau, mask)
        c2 = rgcn_layer_isolated(p, "ent_encoder.layer_2.", cfg, first, relu=True)
        second = attention_mix(p, "ent_encoder.layer_2.", cfg, c2 + p["ent_encoder.layer_2.time_embed"][int(t)],
                               prev2, tau, mask)
        return second if cfg.rec_only_last_layer else torch.max(first, second)

    # ---- dense-history scan (uni / one direction)   models/DynamicRGCN.py:35-54, 156-174 --------
    def _gather_prev(self, graphs, hist, start, cur_t):
        f, s, dt = [], [], []
        for i, g in enumerate(graphs):
            idx = torch.from_numpy(g.ids)
            f.append(hist[i][0][idx])
            s.append(hist[i][1][idx])
            dt.append((cur_t - start[i][idx]).view(-1, 1))
        return torch.cat(f), torch.cat(s), torch.cat(dt)

    def _scan(self, time_batched, direction: str, sample_rate: Optional[float] = None):
        cfg = self.cfg
        L, bsz = cfg.seq_len, len(time_batched[0])
        hist = torch.zeros(bsz, 2, self.M, self.D)
        start = torch.zeros(bsz, self.M)
        for k in range(L - 1):
            ts = [t for t in time_batched[k] if t is not None]
            if not ts:
                continue
            graphs = [self.gd[t] for t in ts]
            if sample_rate is not None:                                    # --random-dropout, DynamicRGCN.py:161,171
                graphs = [sample_edges(g, sample_rate) for g in graphs]
            p1, p2, dt = self._gather_prev(graphs, hist, start, k)
            first, second = self.enc_recurrent(graphs, ts, [p1], [p2], [dt], direction)
            hist = torch.zeros(bsz, 2, self.M, self.D)                   # DynamicRGCN.py:48 (fresh zeros)
            off = 0
            for i, g in enumerate(graphs):
                idx = torch.from_numpy(g.ids)
                hist[i][0][idx] = first[off:off + g.num_nodes]
                hist[i][1][idx] = second[off:off + g.num_nodes]
                start[i][idx] = k
                off += g.num_nodes
        if direction == "backward":                                       # BiDynamicRGCN.py:97-99
            hist, start = torch.flip(hist, [0]), torch.flip(start, [0])
        return hist, start

    # ---- public: what evaluate_embed computes -----------------------------------------------
    def evaluate_embed(self, t_list: Sequence[int]):
        """Region R1 of SURVEY section 8(d).  Returns a dict with 'per_graph' (list of [N_i, D] for the
        target graphs in descending-time order), 'times' (their timestamps) and the history state
        needed by all_embeds()."""
        cfg = self.cfg
        if cfg.module == "SRGCN":                                         # baselines/StaticRGCN.py:23-28
            ts = [int(t) for t in t_list]
            graphs = [self.gd[t] for t in ts]
            out = self.enc_static(graphs, ts)
            return {"per_graph": list(out.split([g.num_nodes for g in graphs])), "times": ts, "graphs": graphs}
        L = cfg.seq_len
        tb_f = window_forward(t_list, L, self.times)
        ts = tb_f[-1]
        graphs = [self.gd[t] for t in ts]
        sizes = [g.num_nodes for g in graphs]
        res = {"times": ts, "graphs": graphs}
        if cfg.attention:
            hist, mask = self._attention_history(t_list)
            res.update(hist=hist, mask=mask)
            tau = self._tau()
            p1, p2, lm = [], [], []
            for i, g in enumerate(graphs):                                 # SelfAttentionRGCN.py:73-84
                idx = torch.from_numpy(g.ids)
                p1.append(hist[:, i, 0][:, idx])
                p2.append(hist[:, i, 1][:, idx])
                lm.append(mask[:, i][:, idx])
            out = self.enc_attention_final(graphs, ts, torch.cat(p1, 1).transpose(0, 1),
                                           torch.cat(p2, 1).transpose(0, 1), tau,
                                           torch.cat(lm, 1).transpose(0, 1))
        elif cfg.bidirectional:                                           # BiDynamicRGCN.py:151-163
            hist_f, start_f = self._scan(tb_f, "forward")
            hist_b, start_b = self._scan(window_backward(t_list, L, self.times), "backward")
            f1, f2, dtf = self._gather_prev(graphs, hist_f, start_f, L - 1)
            b1, b2, dtb = self._gather_prev(graphs, hist_b, start_b, L - 1)
            _, out = self.enc_recurrent(graphs, ts, [f1, b1], [f2, b2], [dtf, dtb], direction=None)
            res.update(hist_f=hist_f, start_f=start_f, hist_b=hist_b, start_b=start_b)
        else:                                                             # DynamicRGCN.py:132-144
            hist, start = self._scan(tb_f, "forward")
            p1, p2, dt = self._gather_prev(graphs, hist, start, L - 1)
            _, out = self.enc_recurrent(graphs, ts, [p1], [p2], [dt], "forward")
            res.update(hist=hist, start=start)
        res["per_graph"] = list(out.split(sizes))
        return res

    def train_loss(self, t_list: Sequence[int], negative_rate: int, num_pos_facts: int,
                   random_dropout: bool = False) -> Tensor:
        """The training forward of models/DynamicRGCN.py:176-194 (GRRGCN / RRGCN) and models/BiDynamicRGCN.py:165-187
        (BiGRRGCN / BiRRGCN) with dropout p = 0: history steps
        on full graphs (or np.random.choice 80 % edge subsets with --random-dropout), the final step on a 50 % edge
        subset with recomputed norms (DynamicRGCN.py:76-94), then per target graph the negative sampler on the FULL
        graph and the tail + head cross-entropy (TKG_Module.py:202-213).  Global NumPy / torch RNG order as in the
        reference."""
        cfg = self.cfg
        L = cfg.seq_len
        rate_hist = 0.8 if random_dropout else None
        if cfg.module == "SRGCN":
            # baselines/StaticRGCN.py:36-47, 60-89: the target graphs in the CALLER's order, each on a 50 % edge subset
            ts = [int(t) for t in t_list]
            full_graphs = [self.gd[t] for t in ts]
            out = self.enc_static([sample_edges(g, 0.5) for g in full_graphs], ts)
            res = {"times": ts, "graphs": full_graphs, "per_graph": list(out.split([g.num_nodes for g in full_graphs]))}
            return self._link_prediction_loss(res, full_graphs, negative_rate, num_pos_facts)
        if cfg.attention:
            # models/SelfAttentionRGCN.py:122-139 (history sub-sampled at 0.8 only with --random-dropout) and
            # models/BiSelfAttentionRGCN.py:48-69 (history always on full graphs); the final step on a 50 % edge subset
            tb = window_forward(t_list, L, self.times)
            ts = tb[-1]
            full_graphs = [self.gd[t] for t in ts]
            hist, mask = self._attention_history(t_list, sample_rate=None if cfg.bidirectional else rate_hist)
            p1, p2, lm = [], [], []
            for i, g in enumerate(full_graphs):
                idx = torch.from_numpy(g.ids)
                p1.append(hist[:, i, 0][:, idx])
                p2.append(hist[:, i, 1][:, idx])
                lm.append(mask[:, i][:, idx])
            graphs = [sample_edges(g, 0.5) for g in full_graphs]
            out = self.enc_attention_final(graphs, ts, torch.cat(p1, 1).transpose(0, 1), torch.cat(p2, 1).transpose(0, 1),
                                           self._tau(), torch.cat(lm, 1).transpose(0, 1))
            res = {"times": ts, "graphs": full_graphs, "hist": hist, "mask": mask,
                   "per_graph": list(out.split([g.num_nodes for g in full_graphs]))}
            return self._link_prediction_loss(res, full_graphs, negative_rate, num_pos_facts)
        tb = window_forward(t_list, L, self.times)
        ts = tb[-1]
        full_graphs = [self.gd[t] for t in ts]
        sizes = [g.num_nodes for g in full_graphs]
        if cfg.bidirectional:                       # models/BiDynamicRGCN.py:165-174: forward scan, backward scan, centre
            hist_f, start_f = self._scan(tb, "forward", sample_rate=rate_hist)
            hist_b, start_b = self._scan(window_backward(t_list, L, self.times), "backward", sample_rate=rate_hist)
            f1, f2, dtf = self._gather_prev(full_graphs, hist_f, start_f, L - 1)
            b1, b2, dtb = self._gather_prev(full_graphs, hist_b, start_b, L - 1)
            graphs = [sample_edges(g, 0.5) for g in full_graphs]
            _, out = self.enc_recurrent(graphs, ts, [f1, b1], [f2, b2], [dtf, dtb], direction=None)
            res = {"times": ts, "graphs": full_graphs, "hist_f": hist_f, "start_f": start_f, "hist_b": hist_b,
                   "start_b": start_b, "per_graph": list(out.split(sizes))}
        else:
            hist, start = self._scan(tb, "forward", sample_rate=rate_hist)
            p1, p2, dt = self._gather_prev(full_graphs, hist, start, L - 1)
            graphs = [sample_edges(g, 0.5) for g in full_graphs]
            _, out = self.enc_recurrent(graphs, ts, [p1], [p2], [dt], "forward")
            res = {"times": ts, "graphs": full_graphs, "hist": hist, "start": start, "per_graph": list(out.split(sizes))}
        return self._link_prediction_loss(res, full_graphs, negative_rate, num_pos_facts)

    def _link_prediction_loss(self, res, full_graphs, negative_rate: int, num_pos_facts: int) -> Tensor:
        """Per target graph: the negative sampler on the FULL graph, then tail + head cross-entropy against the
        all-entity table (models/TKG_Module.py:202-213)."""
        score = {"complex": score_complex, "distmult": score_distmult, "transE": score_transe}["complex"]
        rel = self.p["rel_embeds"]
        loss = torch.zeros(())
        for i, g in enumerate(full_graphs):
            tri, neg_t, neg_h, lab = negative_samples(g, self.M, negative_rate, num_pos_facts)
            tri, neg_t, neg_h, lab = (torch.from_numpy(np.asarray(x)).long() for x in (tri, neg_t, neg_h, lab))
            ent = res["per_graph"][i]
            alls = self.all_embeds(res, i)
            r = rel[tri[:, 1]]
            loss = loss + torch.nn.functional.cross_entropy(score(ent[tri[:, 0]], r, alls[neg_t], mode="tail"), lab)
            loss = loss + torch.nn.functional.cross_entropy(score(alls[neg_h], r, ent[tri[:, 2]], mode="head"), lab)
        return loss

    def _tau(self) -> Tensor:
        L = self.cfg.seq_len
        if self.cfg.bidirectional:                                        # BiSelfAttentionRGCN.py:19-20
            return torch.tensor(list(range(L - 1, 0, -1)) * 2 + [0.0])
        return torch.tensor(list(range(L - 1, -1, -1))).float()           # SelfAttentionRGCN.py:22-23

    def _attention_scan(self, time_batched, flip: bool, sample_rate: Optional[float] = None):
        """models/SelfAttentionRGCN.py:104-120 / BiSelfAttentionRGCN.py:25-46 (sample_rate: the uni model's
        --random-dropout history sub-sampling in train mode, SelfAttentionRGCN.py:111, 118)."""
        L, bsz = self.cfg.seq_len, len(time_batched[0])
        hist = torch.zeros(L - 1, bsz, 2, self.M, self.D)
        mask = torch.zeros(L - 1, bsz, self.M) - 10e9
        for k in range(L - 1):
            ts = [t for t in time_batched[k] if t is not None]
            if not ts:
                continue
            graphs = [self.gd[t] for t in ts]
            if sample_rate is not None:
                graphs = [sample_edges(g, sample_rate) for g in graphs]
            first, second = self.enc_attention_history(graphs, ts)
            off = 0
            for i, g in enumerate(graphs):
                idx = torch.from_numpy(g.ids)
                mask[k][i][idx] = 0
                hist[k][i][0][idx] = first[off:off + g.num_nodes]
                hist[k][i][1][idx] = second[off:off + g.num_nodes]
                off += g.num_nodes
        if flip:
            hist, mask = torch.flip(hist, [1]), torch.flip(mask, [1])
        return hist, mask

    def _attention_history(self, t_list, sample_rate: Optional[float] = None):
        L = self.cfg.seq_len
        hist, mask = self._attention_scan(window_forward(t_list, L, self.times), flip=False, sample_rate=sample_rate)
        if self.cfg.bidirectional:                                        # BiSelfAttentionRGCN.py:82-84
            hb, mb = self._attention_scan(window_backward(t_list, L, self.times), flip=True)
            hist = torch.cat([hist, hb], dim=0)
            mask = torch.cat([mask, mb], dim=0)
        mask = torch.cat([mask, torch.zeros(1, *mask.shape[1:])], dim=0)   # current slot
        return hist, mask

    def all_embeds(self, res, i: int) -> Tensor:
        """get_all_embeds_Gt for batch item i (region R2): forward_isolated over all M entities with
        the item's history, then rows of active entities overwritten by the graph outputs
        (models/DynamicRGCN.py:56-64, BiDynamicRGCN.py:102-112, SelfAttentionRGCN.py:26-43,
        baselines/StaticRGCN.py:48-58)."""
        cfg = self.cfg
        g, t = res["graphs"][i], res["times"][i]
        L = cfg.seq_len
        if cfg.use_embed_for_non_active:       # DynamicRGCN.py:58-59, BiDynamicRGCN.py:105-106, SelfAttentionRGCN.py:31-32, StaticRGCN.py:51-52
            out = self.p["ent_embeds"].clone()
        elif cfg.module == "SRGCN":
            out = self.enc_static_isolated(t).clone()
        elif cfg.attention:
            hist, mask = res["hist"], res["mask"]
            out = self.enc_attention_isolated(t, hist[:, i, 0].transpose(0, 1), hist[:, i, 1].transpose(0, 1),
                                              self._tau(), mask[:, i].transpose(0, 1))
        elif cfg.bidirectional:
            dtf = (L - 1 - res["start_f"][i]).unsqueeze(-1)
            dtb = (L - 1 - res["start_b"][i]).unsqueeze(-1)
            out = self.enc_recurrent_isolated(t, [res["hist_f"][i][0], res["hist_b"][i][0]],
                                              [res["hist_f"][i][1], res["hist_b"][i][1]], [dtf, dtb])
        else:
            dt = (L - 1 - res["start"][i]).unsqueeze(-1)
            out = self.enc_recurrent_isolated(t, [res["hist"][i][0]], [res["hist"][i][1]], [dt])
        out = out.clone()
        out[torch.from_numpy(g.ids)] = res["per_graph"][i]
        return out


    def all_embeds_with_history_of(self, res, i: int, hist_item: int) -> Tensor:
        """get_all_embeds_Gt of target graph i evaluated with the history tensors of batch item ``hist_item`` -- what
        the reference's calc_metrics computes when an earlier evaluation graph of the batch had no edges (see
        ``evaluate``).  Rows of graph i's entities still come from graph i's states, the time embedding from its time."""
        if hist_item == i or self.cfg.module == "SRGCN":
            return self.all_embeds(res, i)
        shifted = dict(res)
        for key in ("hist", "start", "hist_f", "start_f", "hist_b", "start_b", "mask"):
            if key in res:
                v = res[key]
                if self.cfg.attention:                                    # [L-1, B, ...]: item axis is 1
                    v = v.clone()
                    v[:, i] = res[key][:, hist_item]
                else:                                                     # [B, ...]
                    v = v.clone()
                    v[i] = res[key][hist_item]
                shifted[key] = v
        return self.all_embeds(shifted, i)

    def evaluate(self, t_list: Sequence[int], valid, test=None, val: bool = True, score_function: str = "complex"):
        """``evaluate(t_list, val)`` of the reference -> (ranks LongTensor, mean link-classification loss):
        models/DynamicRGCN.py:118-130 + calc_metrics 196-220 (BiDynamicRGCN.py:176-209, SelfAttentionRGCN.py:141-176,
        baselines/StaticRGCN.py:20-22, 91-113) over utils/evaluation.py:6-106.

        Quirk kept: in the dynamic models' calc_metrics the history index ``i`` only advances for evaluation graphs
        that HAVE edges (``continue`` sits before ``i += 1``), so after an empty graph every later graph of the batch
        is scored with the history of the item before it.  StaticRGCN has no history and no such index."""
        cfg = self.cfg
        res = self.evaluate_embed(t_list)
        graph_dict = valid if val else test
        score = {"complex": score_complex, "distmult": score_distmult, "transE": score_transe}[score_function]
        rel = self.p["rel_embeds"]
        ranks, losses = [], []
        i = 0
        for pos, (t, ent) in enumerate(zip(res["times"], res["per_graph"])):
            g = graph_dict[int(t)]
            alls = self.all_embeds(res, pos) if cfg.module == "SRGCN" else self.all_embeds_with_history_of(res, pos, i)
            if g.num_edges == 0:
                continue
            samples = torch.from_numpy(np.stack([g.src, g.rel, g.dst], axis=1)).long()
            parts = [x for x in (self.gd.get(int(t)), valid.get(int(t)), test.get(int(t)) if test is not None else None)
                     if x is not None]
            ranks.append(filtered_ranks(score, ent, rel, alls, samples, g, parts))
            s, r, o = ent[samples[:, 0]], rel[samples[:, 1]], ent[samples[:, 2]]
            losses.append(float(torch.nn.functional.binary_cross_entropy_with_logits(score(s, r, o), torch.ones(samples.shape[0]))))
            i += 1
        ranks = torch.cat(ranks) if ranks else torch.zeros(0, dtype=torch.long)
        return ranks, (float(np.mean(losses)) if losses else float("nan"))


def filtered_ranks(score, ent: Tensor, rel: Tensor, alls: Tensor, samples: Tensor, g: "SnapGraph", splits) -> Tensor:
    """utils/evaluation.py:35-106 for one evaluation graph: true heads / tails from the train + valid + test triples of
    the timestamp in LOCAL node ids (lines 17-33; the three graphs of a timestamp share one node set), per query the
    known answers masked except the query's own target (82-99), scores against all entities in batches of 100,
    torch.where(mask, -10e6, score), sigmoid, descending sort, position of the target (53-80, 101-106); subject side
    first, then object side, + 1 (47-49)."""
    tri = np.concatenate([np.stack([x.src, x.rel, x.dst], axis=1) for x in splits], axis=0)
    th, tt = {}, {}
    for h, r, t in tri.tolist():
        tt.setdefault((h, r), []).append(t)
        th.setdefault((r, t), []).append(h)
    th = {k: np.array(list(set(v))) for k, v in th.items()}
    tt = {k: np.array(list(set(v))) for k, v in tt.items()}
    ids = g.ids
    M = alls.shape[0]
    n = samples.shape[0]

    def mask_of(mode):
        mask = torch.zeros(n, M, dtype=torch.bool)
        for q, (h, r, t) in enumerate(samples.tolist()):
            if mode == "tail":
                mask[q, torch.from_numpy(ids[tt[(h, r)]])] = True
                mask[q, int(ids[t])] = False
            else:
                mask[q, torch.from_numpy(ids[th[(r, t)]])] = True
                mask[q, int(ids[h])] = False
        return mask

    gids = torch.from_numpy(ids)
    out = {}
    for mode in ("tail", "head"):                                         # ranks_o is computed first (lines 44-47)
        mask = mask_of(mode)
        rk = []
        for lo in range(0, n, 100):
            sl = slice(lo, min(n, lo + 100))
            r = rel[samples[sl, 1]]
            if mode == "tail":
                sc = score(ent[samples[sl, 0]], r, alls, mode="tail")
                target = gids[samples[sl, 2]]
            else:
                sc = score(alls, r, ent[samples[sl, 2]], mode="head")
                target = gids[samples[sl, 0]]
            sc = torch.sigmoid(torch.where(mask[sl], -10e6 * torch.ones_like(sc), sc))
            _, order = torch.sort(sc, dim=1, descending=True)
            rk.append(torch.nonzero(order == target.view(-1, 1))[:, 1].view(-1))
        out[mode] = torch.cat(rk)
    return torch.cat([out["head"], out["tail"]]) + 1


# --------------------------------------------------------------------------------------------
# scores and the negative sampler (kept host-side and bit-exact)
# --------------------------------------------------------------------------------------------
def score_complex(s: Tensor, r: Tensor, o: Tensor, mode: str = "single") -> Tensor:
    """utils/scores.py:27-44."""
    re_s, im_s = torch.chunk(s, 2, dim=-1)
    re_r, im_r = torch.chunk(r, 2, dim=-1)
    re_o, im_o = torch.chunk(o, 2, dim=-1)
    if mode == "head":
        a = re_r * re_o + im_r * im_o
        b = re_r * im_o - im_r * re_o
        return (re_s * a.unsqueeze(1) + im_s * b.unsqueeze(1)).sum(-1)
    a = re_s * re_r - im_s * im_r
    b = re_s * im_r + im_s * re_r
    if mode == "tail":
        return (a.unsqueeze(1) * re_o + b.unsqueeze(1) * im_o).sum(-1)
    return (a * re_o + b * im_o).sum(-1)


def score_distmult(s, r, o, mode="single"):
    """utils/scores.py:4-11."""
    if mode == "tail":
        return torch.sum((s * r).unsqueeze(1) * o, dim=-1)
    if mode == "head":
        return torch.sum(s * (r * o).unsqueeze(1), dim=-1)
    return torch.sum(s * r * o, dim=-1)


def score_transe(s, r, o, mode="single"):
    """utils/scores.py:46-55."""
    if mode == "tail":
        x = (s + r).unsqueeze(1) - o
    elif mode == "head":
        x = s + (r - o).unsqueeze(1)
    else:
        x = s + r - o
    return -torch.norm(x, p=1, dim=-1)


def true_heads_tails(g: SnapGraph):
    """utils/CorrptTriples.py:87-106: (h,r) -> array(set(tails)), (r,t) -> array(set(heads)); the
    set -> list -> array order is CPython's set iteration order and is part of the RNG contract."""
    th, tt = {}, {}
    for h, r, t in zip(g.src.tolist(), g.rel.tolist(), g.dst.tolist()):
        tt.setdefault((h, r), []).append(t)
        th.setdefault((r, t), []).append(h)
    th = {k: np.array(list(set(v))) for k, v in th.items()}
    tt = {k: np.array(list(set(v))) for k, v in tt.items()}
    return th, tt


def negative_samples(g: SnapGraph, num_entities: int, negative_rate: int, num_pos_facts: int):
    """utils/CorrptTriples.py:36-85, same RNG call sequence: torch.randperm(E) iff E > num_pos_facts,
    then per positive triple tail corruption rounds followed by head corruption rounds of
    np.random.randint(num_entities, size=negative_rate) filtered by np.in1d(invert=True)."""
    triples = np.stack([g.src, g.rel, g.dst], axis=1)
    P = min(triples.shape[0], num_pos_facts)
    if num_pos_facts < triples.shape[0]:
        perm = torch.randperm(triples.shape[0]).numpy()
        triples = triples[perm[:num_pos_facts]]
    th, tt = true_heads_tails(g)
    neg_tail = np.zeros((P, 1 + negative_rate), dtype=int)
    neg_head = np.zeros((P, 1 + negative_rate), dtype=int)

    def corrupt(true_local):
        forbidden = [int(g.ids[i]) for i in true_local.tolist()]
        got, size = [], 0
        while size < negative_rate:
            cand = np.random.randint(num_entities, size=negative_rate)
            keep = np.isin(cand, forbidden, assume_unique=True, invert=True)
            cand = cand[keep]
            got.append(cand)
            size += cand.size
        return np.concatenate(got)[:negative_rate]

    for i in range(P):
        h, r, t = (int(x) for x in triples[i])
        tail_s = corrupt(tt[(h, r)])
        head_s = corrupt(th[(r, t)])
        neg_tail[i, 0], neg_head[i, 0] = g.ids[t], g.ids[h]
        neg_tail[i, 1:], neg_head[i, 1:] = tail_s, head_s
    return triples, neg_tail, neg_head, np.zeros(P, dtype=int)


# --------------------------------------------------------------------------------------------
# deterministic, platform-independent parameter fill shared by the golden generator and the tests
# --------------------------------------------------------------------------------------------
def fill_values(name: str, shape, scale: float = 0.2) -> np.ndarray:
    """Exact-integer hash -> float32 in [-scale, scale): reproducible on any machine, so golden
    files store outputs only."""
    import zlib
    n = int(np.prod(shape)) if len(shape) else 1
    seed = np.uint64(zlib.crc32(name.encode()))
    idx = np.arange(n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        x = (idx + seed) * np.uint64(6364136223846793005) + np.uint64(1442695040888963407)
        x ^= x >> np.uint64(33)
        x *= np.uint64(0xFF51AFD7ED558CCD)
        x ^= x >> np.uint64(29)
    frac = ((x >> np.uint64(40)).astype(np.float64) / float(1 << 24)) - 0.5
    return (frac * 2.0 * scale).astype(np.float32).reshape(shape)


def param_shapes(cfg: OracleConfig) -> Dict[str, tuple]:
    """state_dict keys/shapes of the reference shells."""
    D, M, R2, T = cfg.embed_size, cfg.num_ents, 2 * cfg.num_rels, cfg.num_times
    si = so = D // cfg.n_bases
    shp = {"ent_embeds": (M, D), "rel_embeds": (R2, D)}
    for li in (1, 2):
        pre = "ent_encoder.layer_%d." % li
        shp[pre + "time_embed"] = (T, D)
        shp[pre + "weight"] = (R2, cfg.n_bases * si * so)
        shp[pre + "loop_weight"] = (D, D)
        if cfg.layer_bias:
            shp[pre + "h_bias"] = (D,)
        if cfg.learnable_lambda:
            shp[pre + "exponential_decay.weight"] = (1, 1)
            shp[pre + "exponential_decay.bias"] = (1,)
        recurrent_here = (li == 2) or (not cfg.rec_only_last_layer)
        if cfg.module in ("GRRGCN", "BiGRRGCN") and recurrent_here:
            rnns = ["rnn"] if cfg.module == "GRRGCN" else ["forward_rnn", "backward_rnn"]
            for rn in rnns:
                if cfg.type1:
                    shp.update({pre + rn + ".weight_ih": (D, D), pre + rn + ".weight_hh": (3 * D, D),
                                pre + rn + ".bias_ih": (D,), pre + rn + ".bias_hh": (3 * D,)})
                else:
                    shp.update({pre + rn + ".weight_ih_l0": (3 * D, D), pre + rn + ".weight_hh_l0": (3 * D, D),
                                pre + rn + ".bias_ih_l0": (3 * D,), pre + rn + ".bias_hh_l0": (3 * D,)})
        if cfg.module == "RRGCN" and recurrent_here:
            shp[pre + "time_weight"] = (D, D)
        if cfg.module == "BiRRGCN" and recurrent_here:
            shp[pre + "time_weight_forward"] = (D, D)
            shp[pre + "time_weight_backward"] = (D, D)
        if cfg.attention and recurrent_here:
            for nm in ("q_linear", "v_linear", "k_linear"):
                shp[pre + nm + ".weight"] = (D, D)
    return shp


def make_params(cfg: OracleConfig, scale: float = 0.2) -> Dict[str, Tensor]:
    return {k: torch.from_numpy(fill_values(k, s, scale)) for k, s in param_shapes(cfg).items()}
