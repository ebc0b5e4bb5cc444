"""GPU: the reference's per-step encoder calls (SURVEY.md section 8b -- ``ent_encoder.forward / forward_isolated /
forward_one_direction`` with the reference's argument lists) on the CUDA path, held to the oracle's statement of the
same calls (``enc_recurrent``, ``enc_recurrent_isolated``, ``enc_static``; models/RRGCN.py:192-217,
models/BiRRGCN.py:210-257, models/RGCN.py:154-164) on the same seeded inputs, and -- driven step by step the way
models/DynamicRGCN.py:156-174 drives them -- to the one-program window forward ``model.encode``.

Tolerance: 1e-4 relative fp32 (BASELINE.json north star)."""
import numpy as np
import pytest
import torch

from tests.helpers import CASE_BY_NAME, oracle_model, product_model, rel_err

pytestmark = pytest.mark.gpu

RTOL = 1e-4

RECURRENT = ["grrgcn_tiny_d128_last", "grrgcn_tiny_d128_full", "grrgcn_tiny_d128_full_note", "rrgcn_tiny_d128_last",
             "rrgcn_tiny_d128_full", "grrgcn_tiny_d32_nb8_type1", "grrgcn_tiny_d128_lambda", "grrgcn_tiny_d200_nb100",
             "grrgcn_icews_d128_L8"]
BIDIRECTIONAL = ["bigrrgcn_tiny_d128_last", "bigrrgcn_tiny_d128_full", "birrgcn_tiny_d128_full", "bigrrgcn_tiny_d200_nb100",
                 "bigrrgcn_icews_d128_L8"]


def _inputs(model, oracle, n_graphs, seed, rows=None):
    """A few training snapshots + seeded dense previous states (a third of the rows without history) and time gaps."""
    from temp_b200.stepwise import batch
    times = sorted(model.graph_dict_train.keys())[1:1 + n_graphs]
    snaps = [model.graph_dict_train[t] for t in times]
    bg = batch(snaps, model)
    n = bg.number_of_nodes() if rows is None else rows
    g = torch.Generator().manual_seed(seed)
    D = model.embed_size

    def state():
        x = torch.randn(n, D, generator=g) * 0.1
        x[torch.rand(n, generator=g) < 0.33] = 0.0
        return x

    dt = lambda: torch.randint(1, 6, (n, 1), generator=g).float()
    return times, snaps, bg, [state() for _ in range(4)], [dt(), dt()], [oracle.gd[t] for t in times]


def _close(got, want):
    err = rel_err(got.detach().cpu().numpy(), want.detach().cpu().numpy())
    assert err < RTOL, "max rel-to-scale err %.3e" % err


@pytest.mark.parametrize("name", RECURRENT)
def test_uni_encoder_step_calls_match_oracle(name):
    case = CASE_BY_NAME[name]
    model, oracle = product_model(case), oracle_model(case)
    enc = model.ent_encoder
    times, snaps, bg, st, dts, ograph = _inputs(model, oracle, 3, 1)
    cu = lambda x: x.cuda()
    first, second = enc.forward(bg, cu(st[0]), cu(st[1]), cu(dts[0]), torch.tensor(times), bg.node_sizes)
    with torch.no_grad():
        of, os_ = oracle.enc_recurrent(ograph, times, [st[0]], [st[1]], [dts[0]], "forward")
    assert first.shape == second.shape == (bg.number_of_nodes(), model.embed_size)
    _close(second, os_)
    _close(first, of)
    if case["module"] == "GRRGCN":                       # the same tensor
        assert first.data_ptr() == second.data_ptr()
    # all-entity rows
    M = model.num_ents
    _, _, _, stm, dtm, _ = _inputs(model, oracle, 1, 2, rows=M)
    got = enc.forward_isolated(model.ent_embeds, cu(stm[0]), cu(stm[1]), cu(dtm[0]), times[1])
    with torch.no_grad():
        want = oracle.enc_recurrent_isolated(times[1], [stm[0]], [stm[1]], [dtm[0]])
    _close(got, want)


@pytest.mark.parametrize("name", BIDIRECTIONAL)
def test_bi_encoder_step_calls_match_oracle(name):
    case = CASE_BY_NAME[name]
    model, oracle = product_model(case), oracle_model(case)
    enc = model.ent_encoder
    times, snaps, bg, st, dts, ograph = _inputs(model, oracle, 3, 3)
    cu = lambda x: x.cuda()
    tt = torch.tensor(times)
    for fwd, direction in ((True, "forward"), (False, "backward")):          # history steps, BiRRGCN.py:228-240
        first, second = enc.forward_one_direction(bg, cu(st[0]), cu(st[1]), cu(dts[0]), tt, bg.node_sizes, fwd)
        with torch.no_grad():
            of, os_ = oracle.enc_recurrent(ograph, times, [st[0]], [st[1]], [dts[0]], direction)
        _close(second, os_)
        _close(first, of)
    got = enc.forward(bg, cu(st[0]), cu(st[1]), cu(dts[0]), cu(st[2]), cu(st[3]), cu(dts[1]), tt, bg.node_sizes)   # centre step
    with torch.no_grad():
        _, want = oracle.enc_recurrent(ograph, times, [st[0], st[2]], [st[1], st[3]], dts, None)
    _close(got, want)
    M = model.num_ents
    _, _, _, stm, dtm, _ = _inputs(model, oracle, 1, 4, rows=M)
    got = enc.forward_isolated(model.ent_embeds, cu(stm[0]), cu(stm[1]), cu(dtm[0]), cu(stm[2]), cu(stm[3]), cu(dtm[1]), times[0])
    with torch.no_grad():
        want = oracle.enc_recurrent_isolated(times[0], [stm[0], stm[2]], [stm[1], stm[3]], dtm)
    _close(got, want)


@pytest.mark.parametrize("name", ["srgcn_tiny_d128", "srgcn_tiny_d32_nb8_note"])
def test_static_encoder_step_calls_match_oracle(name):
    case = CASE_BY_NAME[name]
    model, oracle = product_model(case), oracle_model(case)
    enc = model.ent_encoder
    times, snaps, bg, _, _, ograph = _inputs(model, oracle, 4, 5)
    out = enc.forward(bg, times, bg.node_sizes)
    assert out is not bg and "h" in out.ndata                      # a new graph object, like g.local_var() upstream
    with torch.no_grad():
        _close(out.ndata["h"], oracle.enc_static(ograph, times))
        _close(enc.forward_isolated(model.ent_embeds, times[2]), oracle.enc_static_isolated(times[2]))


@pytest.mark.parametrize("name", ["sargcn_tiny_d128_last", "sargcn_tiny_d128_full", "sargcn_tiny_d128_lambda",
                                  "bisargcn_tiny_d128_last", "bisargcn_tiny_d64_full", "bisargcn_icews_d128_L8"])
def test_attention_encoder_step_calls_match_oracle(name):
    """SARGCN.forward / forward_final / forward_isolated (models/SARGCN.py:103-125) with dense [N, T, D] histories and
    the additive mask of models/SelfAttentionRGCN.py:104-120 (0 active, -10e9 inactive)."""
    case = CASE_BY_NAME[name]
    model, oracle = product_model(case), oracle_model(case)
    enc = model.ent_encoder
    times, snaps, bg, _, _, ograph = _inputs(model, oracle, 3, 6)
    D, M = model.embed_size, model.num_ents
    tau = oracle._tau()
    T = tau.numel() - 1
    g = torch.Generator().manual_seed(9)

    def dense(n):
        mask = torch.where(torch.rand(n, T + 1, generator=g) < 0.4, torch.tensor(-10e9), torch.tensor(0.0))
        mask[:, T] = 0.0                                                   # the current step always attends to itself
        mask[: n // 4, :T] = -10e9                                         # rows without any history
        prev = [torch.randn(n, T, D, generator=g) * 0.1 * (mask[:, :T] == 0).unsqueeze(-1) for _ in range(2)]
        return prev, mask

    cu = lambda x: x.cuda()
    first, second = enc.forward(bg, torch.tensor(times), bg.node_sizes)
    with torch.no_grad():
        of, os_ = oracle.enc_attention_history(ograph, times)
    _close(first, of)
    _close(second, os_)
    n = bg.number_of_nodes()
    prev, mask = dense(n)
    got = enc.forward_final(bg, cu(prev[0]), cu(prev[1]), cu(tau), cu(mask), torch.tensor(times), bg.node_sizes)
    with torch.no_grad():
        want = oracle.enc_attention_final(ograph, times, prev[0], prev[1], tau, mask)
    _close(got, want)
    prev, mask = dense(M)
    got = enc.forward_isolated(model.ent_embeds, cu(prev[0]), cu(prev[1]), cu(tau), cu(mask), times[1])
    with torch.no_grad():
        want = oracle.enc_attention_isolated(times[1], prev[0], prev[1], tau, mask)
    _close(got, want)


@pytest.mark.parametrize("name", ["grrgcn_icews_d128_L8", "grrgcn_tiny_d128_full", "rrgcn_tiny_d128_full"])
def test_window_driven_step_by_step_equals_the_one_program_forward(name):
    """models/DynamicRGCN.py:156-174 restated over the per-step calls: dense [B, 2, M, D] history re-zeroed every step
    ("history forgets"), start times kept, previous rows gathered per graph -- the states of the target graphs must equal
    ``model.encode`` (same kernels; the one-program path only fuses and re-schedules them)."""
    from temp_b200.planner import _window_times
    from temp_b200.stepwise import batch
    case = CASE_BY_NAME[name]
    model = product_model(case)
    enc = model.ent_encoder
    t_list = case["t_list"]
    L, M, D = model.train_seq_len, model.num_ents, model.embed_size
    all_times = sorted(model.graph_dict_train.keys())
    rows = _window_times(t_list, L, all_times, backward=False)            # rows[item][step]
    B = len(rows)
    hist = torch.zeros(B, 2, M, D, device="cuda")
    start = torch.zeros(B, M, device="cuda")
    out = None
    for k in range(L):
        items = [i for i in range(B) if rows[i][k] is not None]
        if not items:
            continue
        snaps = [model.graph_dict_train[rows[i][k]] for i in items]
        bg = batch(snaps, model)
        ids = [torch.from_numpy(s.node_ids).cuda() for s in snaps]
        p1 = torch.cat([hist[i, 0, idx] for i, idx in zip(items, ids)])
        p2 = torch.cat([hist[i, 1, idx] for i, idx in zip(items, ids)])
        dt = torch.cat([(k - start[i, idx]).view(-1, 1) for i, idx in zip(items, ids)])
        first, second = enc.forward(bg, p1, p2, dt, [rows[i][k] for i in items], bg.node_sizes)
        if k == L - 1:
            out = second
            break
        hist = torch.zeros_like(hist)
        for i, idx, f, s in zip(items, ids, first.split(bg.node_sizes), second.split(bg.node_sizes)):
            hist[i, 0, idx], hist[i, 1, idx] = f, s
            start[i, idx] = k
    want = model.encode(t_list).out
    assert rel_err(out.cpu().numpy(), want.cpu().numpy()) < RTOL
