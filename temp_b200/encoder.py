"""Parameter containers of the encoder, named exactly like the reference's modules so that
``state_dict`` keys match (checkpoint compatibility):

  ent_encoder.layer_{1,2}.{time_embed, weight, loop_weight[, h_bias][, exponential_decay.*]}
  GRRGCN   : layer_k.rnn.{weight_ih_l0, weight_hh_l0, bias_ih_l0, bias_hh_l0}   (torch.nn.GRU)
             layer_k.rnn.{weight_ih, weight_hh, bias_ih, bias_hh}               (--type1, GRU_cell.py)
  BiGRRGCN : layer_k.{forward_rnn, backward_rnn}.*
  RRGCN    : layer_k.time_weight         BiRRGCN: layer_k.time_weight_{forward, backward}
  SARGCN   : layer_k.{q_linear, v_linear, k_linear}.weight

These classes only OWN parameters (initialised like the reference: models/RGCN.py:15-44,
models/RRGCN.py:64-75, 120-128, models/BiRRGCN.py:9-25, 102-113, models/SARGCN.py:10-23,
models/GRU_cell.py:7-15).  The arithmetic is done by the CUDA kernels of libtemp_b200.so, driven by
the launch programs built in temp_b200/models.py -- none of these modules has a torch forward.
"""
from __future__ import annotations

import torch
import torch.nn as nn

_GAIN = nn.init.calculate_gain("relu")


class GRUCell(nn.Module):
    """Parameters of the ``--type1`` cell (reference models/GRU_cell.py:7-15)."""

    def __init__(self, input_size, hidden_size):
        super().__init__()
        self.input_size, self.hidden_size = input_size, hidden_size
        self.weight_ih = nn.Parameter(torch.randn(hidden_size, input_size))
        self.weight_hh = nn.Parameter(torch.randn(3 * hidden_size, hidden_size))
        self.bias_ih = nn.Parameter(torch.randn(hidden_size))
        self.bias_hh = nn.Parameter(torch.randn(3 * hidden_size))


def _rnn(args, in_feat, out_feat):
    if getattr(args, "type1", False):
        return GRUCell(in_feat, out_feat)
    return nn.GRU(input_size=in_feat, hidden_size=out_feat, num_layers=getattr(args, "num_layers", 1))


class RGCNLayer(nn.Module):
    def __init__(self, args, in_feat, out_feat, num_rels, num_bases, total_times, bias=True, activation=None,
                 self_loop=False, dropout=0.0):
        super().__init__()
        assert num_bases > 0
        self.bias, self.activation, self.self_loop = bool(bias), activation, self_loop
        self.num_rels, self.num_bases = num_rels, num_bases
        self.in_feat, self.out_feat = in_feat, out_feat
        self.submat_in, self.submat_out = in_feat // num_bases, out_feat // num_bases
        self.time_embed = nn.Parameter(torch.empty(len(total_times), in_feat))
        nn.init.xavier_uniform_(self.time_embed, gain=_GAIN)
        self.weight = nn.Parameter(torch.empty(num_rels, num_bases * self.submat_in * self.submat_out))
        nn.init.xavier_uniform_(self.weight, gain=_GAIN)
        if self.bias:
            self.h_bias = nn.Parameter(torch.zeros(out_feat))
        if self.self_loop:
            self.loop_weight = nn.Parameter(torch.empty(in_feat, out_feat))
            nn.init.xavier_uniform_(self.loop_weight, gain=_GAIN)
        self.dropout_p = float(dropout or 0.0)
        self.inv_temperature = args.inv_temperature
        self.learnable_lambda = bool(getattr(args, "learnable_lambda", False))
        if self.learnable_lambda:
            self.exponential_decay = nn.Linear(1, 1)


class GRRGCNLayer(RGCNLayer):
    def __init__(self, args, in_feat, out_feat, *a, **k):
        super().__init__(args, in_feat, out_feat, *a, **k)
        self.rnn = _rnn(args, in_feat, out_feat)


class RRGCNLayer(RGCNLayer):
    def __init__(self, args, in_feat, out_feat, *a, **k):
        super().__init__(args, in_feat, out_feat, *a, **k)
        self.time_weight = nn.Parameter(torch.empty(in_feat, out_feat))
        nn.init.xavier_uniform_(self.time_weight, gain=_GAIN)


class BiGRRGCNLayer(RGCNLayer):
    def __init__(self, args, in_feat, out_feat, *a, **k):
        super().__init__(args, in_feat, out_feat, *a, **k)
        self.forward_rnn = _rnn(args, in_feat, out_feat)
        self.backward_rnn = _rnn(args, in_feat, out_feat)


class BiRRGCNLayer(RGCNLayer):
    def __init__(self, args, in_feat, out_feat, *a, **k):
        super().__init__(args, in_feat, out_feat, *a, **k)
        self.time_weight_forward = nn.Parameter(torch.empty(in_feat, out_feat))
        nn.init.xavier_uniform_(self.time_weight_forward, gain=_GAIN)
        self.time_weight_backward = nn.Parameter(torch.empty(in_feat, out_feat))
        nn.init.xavier_uniform_(self.time_weight_backward, gain=_GAIN)


class SARGCNLayer(RGCNLayer):
    def __init__(self, args, in_feat, out_feat, *a, **k):
        super().__init__(args, in_feat, out_feat, *a, **k)
        self.q_linear = nn.Linear(in_feat, in_feat, bias=False)
        self.v_linear = nn.Linear(in_feat, in_feat, bias=False)
        self.k_linear = nn.Linear(in_feat, in_feat, bias=False)
        self.h = 8
        self.d_k = in_feat // self.h


class _TwoLayer(nn.Module):
    def __init__(self, args, l1_cls, l2_cls, hidden_size, embed_size, num_rels, total_times, bias, act2):
        super().__init__()
        self.rec_only_last_layer = bool(args.rec_only_last_layer)
        self.use_time_embedding = bool(args.use_time_embedding)
        kw = dict(bias=bias, self_loop=True, dropout=args.dropout)
        self.layer_1 = l1_cls(args, embed_size, hidden_size, 2 * num_rels, args.n_bases, total_times,
                              activation=None, **kw)
        self.layer_2 = l2_cls(args, hidden_size, hidden_size, 2 * num_rels, args.n_bases, total_times,
                              activation=act2, **kw)


class RGCN(_TwoLayer):
    """models/RGCN.py:145-152 (static encoder of SRGCN)."""

    def __init__(self, args, hidden_size, embed_size, num_rels, total_times):
        args.rec_only_last_layer = getattr(args, "rec_only_last_layer", False)
        super().__init__(args, RGCNLayer, RGCNLayer, hidden_size, embed_size, num_rels, total_times, True, "relu")

    # per-step calls of the reference's drivers, on the CUDA path (temp_b200/stepwise.py)
    def forward(self, batched_graph, time_batched_list_t, node_sizes=None):
        """models/RGCN.py:154-159 -> the batched graph, ``ndata['h']`` = layer-2 output (+ time embedding)."""
        from .stepwise import static_graph_step
        return static_graph_step(self, batched_graph, time_batched_list_t)

    def forward_isolated(self, ent_embeds, time):
        """models/RGCN.py:161-164."""
        from .stepwise import static_isolated_step
        return static_isolated_step(self, ent_embeds, time)


class RRGCN(_TwoLayer):
    """models/RRGCN.py:170-190."""

    def __init__(self, args, hidden_size, embed_size, num_rels, total_times):
        rec = {"GRRGCN": GRRGCNLayer, "RRGCN": RRGCNLayer}[args.module]
        l1 = RGCNLayer if args.rec_only_last_layer else rec
        super().__init__(args, l1, rec, hidden_size, embed_size, num_rels, total_times, False, None)
        self.impute = bool(getattr(args, "impute", False))
        if self.impute:
            self.impute_weight = nn.Linear(1, 1)

    # per-step calls of the reference's drivers, on the CUDA path (temp_b200/stepwise.py)
    def forward(self, batched_graph, first_prev_graph_embeds, second_prev_graph_embeds, time_diff_tensor,
                time_batched_list_t, node_sizes=None):
        """models/RRGCN.py:192-204 -> (first, second) node states of the batched graph."""
        from .stepwise import graph_step
        return graph_step(self, batched_graph, time_batched_list_t, [first_prev_graph_embeds], [second_prev_graph_embeds],
                          [time_diff_tensor], ["f"])

    def forward_isolated(self, ent_embeds, first_prev_graph_embeds, second_prev_graph_embeds, time_diff_tensor, time):
        """models/RRGCN.py:206-217 -> second-layer states of all rows of ``ent_embeds``."""
        from .stepwise import isolated_step
        return isolated_step(self, ent_embeds, time, [first_prev_graph_embeds], [second_prev_graph_embeds],
                             [time_diff_tensor], ["f"])

    # post-ensemble / impute calls (the Impute* / PostEnsemble* drivers of models/PostDynamicRGCN.py; GRU flavour)
    def forward_post_ensemble(self, batched_graph, first_prev_graph_embeds, second_prev_graph_embeds, time_diff_tensor,
                              time_batched_list_t, node_sizes=None):
        """models/RRGCN.py:219-234 -> (second_local, first, second): the layer-2 output BEFORE the cell ("local"
        stream) next to the recurrent states, every one with the layer-2 time embedding."""
        from .stepwise import graph_step
        return graph_step(self, batched_graph, time_batched_list_t, [first_prev_graph_embeds], [second_prev_graph_embeds],
                          [time_diff_tensor], ["f"], want_local=True)

    def forward_post_ensemble_isolated(self, ent_embeds, first_prev_graph_embeds, second_prev_graph_embeds, time_diff_tensor,
                                       time, pre_embeds_loc):
        """models/RRGCN.py:236-254 -> (second_local, second) over all rows of ``ent_embeds``."""
        from .stepwise import isolated_step
        return isolated_step(self, ent_embeds, time, [first_prev_graph_embeds], [second_prev_graph_embeds],
                             [time_diff_tensor], ["f"], mode="post", locs=[pre_embeds_loc])

    def forward_isolated_impute(self, ent_embeds, first_prev_graph_embeds, second_prev_graph_embeds, time_diff_tensor, time,
                                pre_embeds_loc):
        """models/RRGCN.py:256-269."""
        from .stepwise import isolated_step
        return isolated_step(self, ent_embeds, time, [first_prev_graph_embeds], [second_prev_graph_embeds],
                             [time_diff_tensor], ["f"], mode="impute", locs=[pre_embeds_loc])

    def calc_impute_weight(self, time_diff_tensor):
        """models/RRGCN.py:271-272."""
        return torch.exp(-torch.clamp(self.impute_weight(time_diff_tensor), min=0))


class BiRRGCN(_TwoLayer):
    """models/BiRRGCN.py:188-208 (layer 2 uses relu, unlike RRGCN)."""

    def __init__(self, args, hidden_size, embed_size, num_rels, total_times):
        rec = {"BiGRRGCN": BiGRRGCNLayer, "BiRRGCN": BiRRGCNLayer}[args.module]
        l1 = RGCNLayer if args.rec_only_last_layer else rec
        super().__init__(args, l1, rec, hidden_size, embed_size, num_rels, total_times, False, "relu")
        self.impute = bool(getattr(args, "impute", False))
        if self.impute:
            self.impute_weight_forward = nn.Linear(1, 1)
            self.impute_weight_backward = nn.Linear(1, 1)

    # per-step calls of the reference's drivers, on the CUDA path (temp_b200/stepwise.py)
    def forward(self, batched_graph, first_prev_graph_embeds_forward, second_prev_graph_embeds_forward,
                time_diff_tensor_forward, first_prev_graph_embeds_backward, second_prev_graph_embeds_backward,
                time_diff_tensor_backward, time_batched_list_t, node_sizes=None):
        """models/BiRRGCN.py:210-226 (the centre step: both directions' cells on one RGCN pass) -> second."""
        from .stepwise import graph_step
        return graph_step(self, batched_graph, time_batched_list_t,
                          [first_prev_graph_embeds_forward, first_prev_graph_embeds_backward],
                          [second_prev_graph_embeds_forward, second_prev_graph_embeds_backward],
                          [time_diff_tensor_forward, time_diff_tensor_backward], ["f", "b"])[1]

    def forward_one_direction(self, batched_graph, first_prev_graph_embeds, second_prev_graph_embeds, time_diff_tensor,
                              time_batched_list_t, node_sizes=None, forward=True):
        """models/BiRRGCN.py:228-240 (history steps) -> (first, second)."""
        from .stepwise import graph_step
        return graph_step(self, batched_graph, time_batched_list_t, [first_prev_graph_embeds], [second_prev_graph_embeds],
                          [time_diff_tensor], ["f" if forward else "b"])

    def forward_isolated(self, ent_embeds, first_prev_graph_embeds_forward, second_prev_graph_embeds_forward,
                         time_diff_tensor_forward, first_prev_graph_embeds_backward, second_prev_graph_embeds_backward,
                         time_diff_tensor_backward, time):
        """models/BiRRGCN.py:242-257."""
        from .stepwise import isolated_step
        return isolated_step(self, ent_embeds, time,
                             [first_prev_graph_embeds_forward, first_prev_graph_embeds_backward],
                             [second_prev_graph_embeds_forward, second_prev_graph_embeds_backward],
                             [time_diff_tensor_forward, time_diff_tensor_backward], ["f", "b"])

    # post-ensemble / impute calls (the Impute* / PostEnsemble* drivers of models/PostBiDynamicRGCN.py; GRU flavour)
    def forward_post_ensemble(self, batched_graph, first_prev_graph_embeds_forward, second_prev_graph_embeds_forward,
                              time_diff_tensor_forward, first_prev_graph_embeds_backward, second_prev_graph_embeds_backward,
                              time_diff_tensor_backward, time_batched_list_t, node_sizes=None):
        """models/BiRRGCN.py:259-275 (centre step) -> (second_local, second)."""
        from .stepwise import graph_step
        local, _, second = graph_step(self, batched_graph, time_batched_list_t,
                                      [first_prev_graph_embeds_forward, first_prev_graph_embeds_backward],
                                      [second_prev_graph_embeds_forward, second_prev_graph_embeds_backward],
                                      [time_diff_tensor_forward, time_diff_tensor_backward], ["f", "b"], want_local=True)
        return local, second

    def forward_post_ensemble_one_direction(self, batched_graph, first_prev_graph_embeds, second_prev_graph_embeds,
                                            time_diff_tensor, time_batched_list_t, node_sizes=None, forward=True):
        """models/BiRRGCN.py:277-293 (history steps) -> (second_local, first, second)."""
        from .stepwise import graph_step
        return graph_step(self, batched_graph, time_batched_list_t, [first_prev_graph_embeds], [second_prev_graph_embeds],
                          [time_diff_tensor], ["f" if forward else "b"], want_local=True)

    def forward_post_ensemble_isolated(self, ent_embeds, first_prev_graph_embeds_forward, second_prev_graph_embeds_forward,
                                       time_diff_tensor_forward, first_prev_graph_embeds_backward,
                                       second_prev_graph_embeds_backward, time_diff_tensor_backward, time,
                                       second_embeds_forward_loc, second_embeds_backward_loc):
        """models/BiRRGCN.py:295-319 -> (second_local, second)."""
        from .stepwise import isolated_step
        return isolated_step(self, ent_embeds, time,
                             [first_prev_graph_embeds_forward, first_prev_graph_embeds_backward],
                             [second_prev_graph_embeds_forward, second_prev_graph_embeds_backward],
                             [time_diff_tensor_forward, time_diff_tensor_backward], ["f", "b"], mode="post",
                             locs=[second_embeds_forward_loc, second_embeds_backward_loc])

    def forward_isolated_impute(self, ent_embeds, first_prev_graph_embeds_forward, second_prev_graph_embeds_forward,
                                time_diff_tensor_forward, first_prev_graph_embeds_backward, second_prev_graph_embeds_backward,
                                time_diff_tensor_backward, time, second_embeds_forward_loc, second_embeds_backward_loc):
        """models/BiRRGCN.py:321-338."""
        from .stepwise import isolated_step
        return isolated_step(self, ent_embeds, time,
                             [first_prev_graph_embeds_forward, first_prev_graph_embeds_backward],
                             [second_prev_graph_embeds_forward, second_prev_graph_embeds_backward],
                             [time_diff_tensor_forward, time_diff_tensor_backward], ["f", "b"], mode="impute",
                             locs=[second_embeds_forward_loc, second_embeds_backward_loc])


class SARGCN(_TwoLayer):
    """models/SARGCN.py:86-101 (forces use_time_embedding, line 92)."""

    def __init__(self, args, hidden_size, embed_size, num_rels, total_time):
        args.use_time_embedding = True
        l1 = RGCNLayer if args.rec_only_last_layer else SARGCNLayer
        super().__init__(args, l1, SARGCNLayer, hidden_size, embed_size, num_rels, total_time, True, "relu")

    # per-step calls of the reference's drivers, on the CUDA path (temp_b200/stepwise.py)
    def forward(self, batched_graph, time_batched_list_t, node_sizes=None):
        """models/SARGCN.py:103-107 (history steps) -> (first + te, second + te)."""
        from .stepwise import attention_history_step
        return attention_history_step(self, batched_graph, time_batched_list_t)

    def forward_final(self, batched_graph, first_layer_prev_embeddings, second_layer_prev_embeddings, time_diff,
                      local_attn_mask, time_batched_list_t, node_sizes=None):
        """models/SARGCN.py:109-117: attention of the current step over the dense history ``[N, T, D]`` under the
        additive mask ``[N, T + 1]``."""
        from .stepwise import attention_final_step
        return attention_final_step(self, batched_graph, first_layer_prev_embeddings, second_layer_prev_embeddings,
                                    time_diff, local_attn_mask, time_batched_list_t)

    def forward_isolated(self, ent_embeds, first_layer_prev_embeddings, second_layer_prev_embeddings, time_diff,
                         local_attn_mask, time):
        """models/SARGCN.py:119-125."""
        from .stepwise import attention_isolated_step
        return attention_isolated_step(self, ent_embeds, first_layer_prev_embeddings, second_layer_prev_embeddings,
                                       time_diff, local_attn_mask, time)

    def forward_post_ensemble(self, batched_graph, second_layer_prev_embeddings, time_diff, local_attn_mask, time_batched_list_t,
                              node_sizes=None):
        """models/SARGCN.py:137-141 (the --post-aggregation return of the layer, SARGCN.py:44-45) -> (second_local, attention)."""
        from .stepwise import attention_post_ensemble_step
        return attention_post_ensemble_step(self, batched_graph, second_layer_prev_embeddings, time_diff, local_attn_mask,
                                            time_batched_list_t)

    def forward_isolated_post_ensemble(self, ent_embeds, second_layer_prev_embeddings, time_diff, local_attn_mask, time):
        """models/SARGCN.py:143-146 -> (second_local, attention) over all rows of ``ent_embeds``."""
        from .stepwise import attention_isolated_post_ensemble_step
        return attention_isolated_post_ensemble_step(self, ent_embeds, second_layer_prev_embeddings, time_diff, local_attn_mask, time)
