"""GPU: the post-ensemble / impute variants on the CUDA path --
the encoder calls (models/RRGCN.py:219-272, models/BiRRGCN.py:259-338, models/SARGCN.py:137-146) against the oracle on seeded
dense inputs, and the Impute* / PostEnsemble* model shells (models/PostDynamicRGCN.py, models/PostBiDynamicRGCN.py) against
the outputs of the unmodified reference (tests/golden/post_*.npz, impute_*.npz).  Tolerance: 1e-4 relative fp32."""
import numpy as np
import pytest
import torch

from tests.golden.cases import POST_CASES
from tests.helpers import CASE_BY_NAME, load_golden, oracle_model, product_model, rel_err
from tests.test_oracle_golden import _post_oracle

pytestmark = pytest.mark.gpu
RTOL = 1e-4


def _close(got, want):
    got = got.detach().cpu().numpy() if torch.is_tensor(got) else np.asarray(got)
    want = want.detach().cpu().numpy() if torch.is_tensor(want) else np.asarray(want)
    err = rel_err(got, want)
    assert err < RTOL, "max rel-to-scale err %.3e" % err


def _dense_inputs(model, n, seed, k):
    g = torch.Generator().manual_seed(seed)
    D = model.embed_size

    def state():
        x = torch.randn(n, D, generator=g) * 0.1
        x[torch.rand(n, generator=g) < 0.33] = 0.0
        return x
    return [state() for _ in range(k)], [torch.randint(1, 6, (n, 1), generator=g).float() for _ in range(2)]


@pytest.mark.parametrize("pc", POST_CASES[:6], ids=lambda c: c["name"])
def test_post_ensemble_and_impute_encoder_calls_match_oracle(pc):
    from temp_b200.stepwise import batch
    case, oracle = _post_oracle(pc)
    model = product_model(case, impute=pc["impute"], post_ensemble=pc["post_ensemble"])
    enc = model.ent_encoder
    bi = model.bidirectional
    times = sorted(model.graph_dict_train.keys())[1:4]
    bg = batch([model.graph_dict_train[t] for t in times], model)
    ograph = [oracle.gd[t] for t in times]
    n, M = bg.number_of_nodes(), model.num_ents
    cu = lambda x: x.cuda()
    with torch.no_grad():
        # ---- graph rows -------------------------------------------------------------------------------------------
        st, dts = _dense_inputs(model, n, 3, 4)
        if not bi:
            loc, first, second = enc.forward_post_ensemble(bg, cu(st[0]), cu(st[1]), cu(dts[0]), torch.tensor(times), bg.node_sizes)
            ol, of, os_ = oracle.enc_post_ensemble(ograph, times, [st[0]], [st[1]], [dts[0]], "forward")
            _close(first, of)
        else:
            loc, second = enc.forward_post_ensemble(bg, cu(st[0]), cu(st[1]), cu(dts[0]), cu(st[2]), cu(st[3]), cu(dts[1]),
                                                    torch.tensor(times), bg.node_sizes)
            ol, _, os_ = oracle.enc_post_ensemble(ograph, times, [st[0], st[2]], [st[1], st[3]], dts, None)
            l1, f1, s1 = enc.forward_post_ensemble_one_direction(bg, cu(st[0]), cu(st[1]), cu(dts[0]), torch.tensor(times),
                                                                 bg.node_sizes, forward=False)
            o1 = oracle.enc_post_ensemble(ograph, times, [st[0]], [st[1]], [dts[0]], "backward")
            for a, b in zip((l1, f1, s1), o1):
                _close(a, b)
        _close(loc, ol)
        _close(second, os_)
        # ---- all-entity rows --------------------------------------------------------------------------------------
        st, dts = _dense_inputs(model, M, 5, 6)
        t = times[1]
        if not bi:
            args = (model.ent_embeds, cu(st[0]), cu(st[1]), cu(dts[0]), t, cu(st[4]))
            o_args = (t, [st[0]], [st[1]], [dts[0]], [st[4]])
        else:
            args = (model.ent_embeds, cu(st[0]), cu(st[1]), cu(dts[0]), cu(st[2]), cu(st[3]), cu(dts[1]), t, cu(st[4]), cu(st[5]))
            o_args = (t, [st[0], st[2]], [st[1], st[3]], dts, [st[4], st[5]])
        gl, gr = enc.forward_post_ensemble_isolated(*args)
        wl, wr = oracle.enc_post_ensemble_isolated(*o_args)
        _close(gl, wl)
        _close(gr, wr)
        if pc["impute"]:
            _close(enc.forward_isolated_impute(*args), oracle.enc_isolated_impute(*o_args))


@pytest.mark.parametrize("pc", POST_CASES, ids=lambda c: c["name"])
def test_post_ensemble_and_impute_models_match_reference(pc):
    """evaluate_embed / get_all_embeds_Gt / evaluate of the shells against what the unmodified reference classes returned."""
    case = CASE_BY_NAME[pc["base"]]
    gold = load_golden(pc["name"])
    model = product_model(case, impute=pc["impute"], post_ensemble=pc["post_ensemble"])
    L, bi = case["L"], model.bidirectional
    ev = model.evaluate_embed(torch.tensor(case["t_list"]), val=True)
    rows = gold["all_rows"]
    if pc["post_ensemble"]:
        loc, rec, graphs = ev[0], ev[1], ev[2]
        hist = ev[4:]
        tl = ev[3][-1] if not bi else ev[3]
        _close(torch.cat(loc), gold["per_graph_loc"])
    else:
        rec, graphs = ev[0], ev[1]
        hist = ev[3:]
        tl = ev[2][-1] if not bi else ev[2]
    assert [int(t) for t in tl] == gold["times"].tolist()
    _close(torch.cat(rec), gold["per_graph"])
    alls = []
    for i in range(len(tl)):
        dense = [hist[0:3]] + ([hist[3:6]] if bi else [])
        a = []
        for l_, r_, s_ in dense:
            a += [l_[i], r_[i][0], r_[i][1], L - 1 - s_[i]]
        pre = (loc[i], rec[i]) if pc["post_ensemble"] else (rec[i],)
        alls.append(model.get_all_embeds_Gt(*pre, graphs[i], tl[i], *a))
    if pc["post_ensemble"]:
        _close(torch.stack([a[0] for a in alls])[:, rows], gold["all_embeds_loc"])
        _close(torch.stack([a[1] for a in alls])[:, rows], gold["all_embeds"])
    else:
        _close(torch.stack(alls)[:, rows], gold["all_embeds"])
    ranks, _ = model.evaluate(torch.tensor(case["t_list"]), val=True)
    want = torch.from_numpy(gold["ranks"]).cuda()
    assert ranks.shape == want.shape
    # (the all-entity tables above agree to 1e-4; a rank moves by one place where two sigmoids of the weighted score sum are
    # within fp32 rounding of each other -- measured: 1.3 % of the 156 ensemble ranks of the real-ICEWS14 case, none elsewhere)
    same = (ranks == want).float().mean().item()
    assert same >= 0.98 and int((ranks - want).abs().max()) <= 1, (same, int((ranks - want).abs().max()))
    with torch.no_grad():
        assert torch.isfinite(model.eval().forward(torch.tensor(case["t_list"])))


@pytest.mark.parametrize("name", ["sargcn_tiny_d128_last", "bisargcn_tiny_d128_last"])
def test_attention_post_ensemble_calls_match_oracle(name):
    """SARGCN.forward_post_ensemble / forward_isolated_post_ensemble (models/SARGCN.py:137-146): layer-2 attention over the
    dense history + the layer-2 output before it ("local", with its time embedding)."""
    from oracle.temp_oracle import attention_mix, batch_graphs, rgcn_layer_graph, rgcn_layer_isolated, time_rows
    from temp_b200.stepwise import batch
    case = CASE_BY_NAME[name]
    model, oracle = product_model(case), oracle_model(case)
    enc = model.ent_encoder
    cfg, p = oracle.cfg, oracle.p
    times = sorted(model.graph_dict_train.keys())[2:5]
    bg = batch([model.graph_dict_train[t] for t in times], model)
    ograph = [oracle.gd[t] for t in times]
    g = torch.Generator().manual_seed(9)
    T = (case["L"] - 1) * (2 if model.bidirectional else 1)
    tau = torch.tensor(model.time_diff(model.plan(case["t_list"])), dtype=torch.float32)

    def inputs(n):
        prev = torch.randn(n, T, model.embed_size, generator=g) * 0.1
        mask = torch.where(torch.rand(n, T + 1, generator=g) < 0.4, torch.tensor(-10e9), torch.tensor(0.0))
        mask[:, -1] = 0.0
        return prev, mask
    with torch.no_grad():
        n = bg.number_of_nodes()
        prev, mask = inputs(n)
        loc, att = enc.forward_post_ensemble(bg, prev.cuda(), tau.cuda(), mask.cuda(), torch.tensor(times), bg.node_sizes)
        obg = batch_graphs(ograph)
        sizes = [x.num_nodes for x in ograph]
        h1 = rgcn_layer_graph(p, "ent_encoder.layer_1.", cfg, oracle._embed(obg), obg, relu=False)
        cur = rgcn_layer_graph(p, "ent_encoder.layer_2.", cfg, h1, obg, relu=True) + time_rows(p, "ent_encoder.layer_2.", times, sizes)
        _close(loc, cur)
        _close(att, attention_mix(p, "ent_encoder.layer_2.", cfg, cur, prev, tau, mask))
        M = model.num_ents
        prev, mask = inputs(M)
        t = times[1]
        loc, att = enc.forward_isolated_post_ensemble(model.ent_embeds, prev.cuda(), tau.cuda(), mask.cuda(), t)
        first = rgcn_layer_isolated(p, "ent_encoder.layer_1.", cfg, p["ent_embeds"], relu=False)
        cur = rgcn_layer_isolated(p, "ent_encoder.layer_2.", cfg, first, relu=True) + p["ent_encoder.layer_2.time_embed"][t]
        _close(loc, cur)
        _close(att, attention_mix(p, "ent_encoder.layer_2.", cfg, cur, prev, tau, mask))
