"""Decoders of the reference (utils/scores.py:4-55) in torch (used by evaluation / link classification) and the fused
training-time scorer on the CUDA path: candidate gather + score + cross-entropy in one
kernel behind ``temp_score_loss_fwd`` -- the [P, 1 + negatives, D] gather of models/TKG_Module.py:202-213 is never
materialised -- with its backward ``temp_score_loss_bwd`` for training."""
import ctypes as C

import torch
import torch.nn.functional as F


def _per_positive_loss(ent_embed, rel_embeds, table, triplets, cand, score_function, corrupt_tail):
    from . import lib
    P, n_cand = int(cand.shape[0]), int(cand.shape[1])
    D = int(table.shape[1])
    assert triplets.dtype == torch.int64 and cand.dtype == torch.int64 and table.is_cuda
    loss = torch.empty(P, dtype=torch.float32, device=table.device)
    a = lib.ScoreLossArgs(P, n_cand, D, lib.SCORE_FN[score_function], int(bool(corrupt_tail)), ent_embed.data_ptr(),
                          rel_embeds.data_ptr(), table.data_ptr(), triplets.data_ptr(), cand.data_ptr(), loss.data_ptr())
    lib.check(lib.load().temp_score_loss_fwd(C.byref(a), C.c_void_p(lib.current_stream())), "temp_score_loss_fwd")
    return loss


class _FusedLinkPredictionLoss(torch.autograd.Function):
    """The fused scorer under autograd: forward ``temp_score_loss_fwd``, backward ``temp_score_loss_bwd`` (scores
    recomputed, gradients of the candidate rows added with vector atomics) -- neither direction materialises the
    [P, 1 + neg, D] gather the reference back-propagates through (models/TKG_Module.py:202-213)."""

    @staticmethod
    def forward(ctx, ent_embed, rel_embeds, table, triplets, cand, score_function, corrupt_tail):
        ent_embed, rel_embeds, table = ent_embed.contiguous(), rel_embeds.contiguous(), table.contiguous()
        triplets, cand = triplets.contiguous(), cand.contiguous()
        ctx.save_for_backward(ent_embed, rel_embeds, table, triplets, cand)
        ctx.score_function, ctx.corrupt_tail = score_function, bool(corrupt_tail)
        return _per_positive_loss(ent_embed, rel_embeds, table, triplets, cand, score_function, corrupt_tail).mean()

    @staticmethod
    def backward(ctx, grad_out):
        from . import lib
        ent_embed, rel_embeds, table, triplets, cand = ctx.saved_tensors
        P, n_cand, D = int(cand.shape[0]), int(cand.shape[1]), int(table.shape[1])
        g = (grad_out.detach().to(torch.float32) / P).expand(P).contiguous()
        g_ent, g_rel, g_table = torch.zeros_like(ent_embed), torch.zeros_like(rel_embeds), torch.zeros_like(table)
        a = lib.ScoreLossBwdArgs(P, n_cand, D, lib.SCORE_FN[ctx.score_function], int(ctx.corrupt_tail), ent_embed.data_ptr(),
                                 rel_embeds.data_ptr(), table.data_ptr(), triplets.data_ptr(), cand.data_ptr(), g.data_ptr(),
                                 g_ent.data_ptr(), g_rel.data_ptr(), g_table.data_ptr())
        lib.check(lib.load().temp_score_loss_bwd(C.byref(a), C.c_void_p(lib.current_stream())), "temp_score_loss_bwd")
        return g_ent, g_rel, g_table, None, None, None, None


def pad_channels(x, score_function="complex"):
    """Zero-pads the channel axis of ``x [n, d]`` to the next multiple of 32 -- the width the scorer / ranking kernels split
    over the lanes of a warp (e.g. d = 200, the reference's n_bases = 100 configuration, -> 224).  ComplEx keeps its
    (real | imaginary) halves: each half is padded on its own.  Zero channels add nothing to any of the three scores
    (bilinear products, or |0 + 0 - 0| for TransE), and the padding is a differentiable torch op."""
    d = int(x.shape[-1])
    if d % 32 == 0:
        return x
    d_pad = (d + 31) // 32 * 32
    if score_function == "complex":
        h, hp = d // 2, d_pad // 2
        return torch.cat([F.pad(x[..., :h], (0, hp - h)), F.pad(x[..., h:], (0, hp - h))], dim=-1)
    return F.pad(x, (0, d_pad - d))


def fused_link_prediction_loss(ent_embed, rel_embeds, triplets, cand, table, score_function="complex", corrupt_tail=True):
    """mean_p [ logsumexp_c score(p, c) - score(p, 0) ]  ==  F.cross_entropy(calc_score(...), zeros)  of
    TKG_Module.train_link_prediction.  ``triplets`` [P, 3] and ``cand`` [P, 1 + neg] are int64 CUDA tensors.
    Differentiable w.r.t. ``ent_embed``, ``rel_embeds`` and ``table`` when gradients are enabled."""
    if int(cand.shape[0]) == 0:
        return table.new_zeros(())
    if int(table.shape[-1]) % 32 != 0:
        ent_embed, rel_embeds, table = (pad_channels(x, score_function) for x in (ent_embed, rel_embeds, table))
    if torch.is_grad_enabled() and (ent_embed.requires_grad or rel_embeds.requires_grad or table.requires_grad):
        return _FusedLinkPredictionLoss.apply(ent_embed, rel_embeds, table, triplets, cand, score_function, corrupt_tail)
    return _per_positive_loss(ent_embed.detach().contiguous(), rel_embeds.detach().contiguous(), table.detach().contiguous(),
                              triplets.contiguous(), cand.contiguous(), score_function, corrupt_tail).mean()


def distmult(s, r, o, mode="single"):
    if mode == "tail":
        return torch.sum((s * r).unsqueeze(1) * o, dim=-1)
    if mode == "head":
        return torch.sum(s * (r * o).unsqueeze(1), dim=-1)
    return torch.sum(s * r * o, dim=-1)


def complex_score(head, relation, tail, mode="single"):
    re_h, im_h = torch.chunk(head, 2, dim=-1)
    re_r, im_r = torch.chunk(relation, 2, dim=-1)
    re_t, im_t = torch.chunk(tail, 2, dim=-1)
    if mode == "head":
        re_s = re_r * re_t + im_r * im_t
        im_s = re_r * im_t - im_r * re_t
        return (re_h * re_s.unsqueeze(1) + im_h * im_s.unsqueeze(1)).sum(dim=-1)
    re_s = re_h * re_r - im_h * im_r
    im_s = re_h * im_r + im_h * re_r
    if mode == "tail":
        return (re_s.unsqueeze(1) * re_t + im_s.unsqueeze(1) * im_t).sum(dim=-1)
    return (re_s * re_t + im_s * im_t).sum(dim=-1)


def transE(head, relation, tail, mode="single"):
    if mode == "tail":
        score = (head + relation).unsqueeze(1) - tail
    elif mode == "head":
        score = head + (relation - tail).unsqueeze(1)
    else:
        score = head + relation - tail
    return -torch.norm(score, p=1, dim=-1)
