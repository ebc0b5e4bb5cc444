"""Decoders of the reference (utils/scores.py:4-55).  Scoring is a "next" row of the scope table
(SURVEY.md section 8f rank 2): it stays in torch for now and is not part of the timed encoder region."""
import torch


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
