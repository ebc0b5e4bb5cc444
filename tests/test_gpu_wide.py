"""GPU: the 64-row tcgen05 layer of temp_b200/csrc/tc_wide.cu (d != 128, and d == 128 with 2x2 / 4x4 relation blocks) called
directly through the C ABI (temp_rgcn_layer_fwd) on random CSR graphs, held to
  (1) an fp64 statement of the layer (models/RGCN.py:53-104 + the chained projection of models/RRGCN.py:84) and
  (2) the fp32 SIMT kernel on the same arguments (the launch without packed operand images),
for BASELINE config 3's shape (d = 200, n_bases = 100 => 2x2 blocks) and the widths around it.  The model-level goldens of the
reference for these widths run in tests/test_gpu_parity.py."""
import ctypes as C
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _log(rec):
    try:
        os.makedirs("gpurun_out", exist_ok=True)
        with open("gpurun_out/wide_layer_stats.jsonl", "a") as f:
            f.write(json.dumps(rec) + "\n")
    except OSError:
        pass


def _problem(D, S, R, row0, chain_n, seed, heavy_at=8, max_deg=40, n_rel=12, T=5, n_src=None, n_heavy=3):
    """Random packed rows [0, R) with a CSR by destination; the launch covers [row0, R)."""
    g = np.random.default_rng(seed)
    n_src = n_src or R
    deg = g.integers(0, 7, size=R)
    deg[g.random(R) < 0.3] = 0
    if n_heavy:
        deg[g.integers(0, R, size=n_heavy)] = g.integers(heavy_at, max_deg, size=n_heavy)      # a few high in-degree rows
    deg[:row0] = 0
    row_ptr = np.zeros(R + 1, np.int32)
    row_ptr[1:] = np.cumsum(deg)
    E = int(row_ptr[-1])
    e_src = g.integers(0, n_src, size=E).astype(np.int32)
    e_rel = g.integers(0, n_rel, size=E).astype(np.int32)
    norm = np.where(deg > 0, 1.0 / np.maximum(deg, 1), 0.0).astype(np.float32)
    rows = np.arange(R)
    has = deg > 0
    light = rows[has & (deg < heavy_at)]
    heavy = rows[has & (deg >= heavy_at)]
    lists = {k: np.stack([v, row_ptr[v], row_ptr[v + 1]], 1).astype(np.int32).reshape(-1, 3) for k, v in (("rows", light), ("heavy", heavy))}
    f = lambda *s, scale=0.5: (g.standard_normal(s) * scale).astype(np.float32)
    p = dict(D=D, S=S, R=R, row0=row0, chain_n=chain_n, row_ptr=row_ptr, e_src=e_src, e_rel=e_rel, norm=norm, lists=lists,
             x=f(n_src, D), weight=f(n_rel, D * S), loop_w=f(D, D, scale=0.5 / np.sqrt(D)), bias=f(D), te=f(T, D),
             row_time=np.sort(g.integers(0, T, size=R)).astype(np.int32), a_index=g.permutation(n_src)[:R].astype(np.int32) if n_src != R else None)
    if chain_n:
        p["chain_w"] = f(D, chain_n, scale=0.5 / np.sqrt(D))
        p["chain_b"] = f(chain_n)
    return p


def _reference(p, act, residual, te_out, te_chain, graph=True):
    """fp64: agg_v = norm_v * sum_e norm_v * blockdiag(W[rel_e]) x[src_e]; out = act(agg (or x) + x W_loop + b); chain."""
    D, S, R = p["D"], p["S"], p["R"]
    x = p["x"].astype(np.float64)
    own = x[p["a_index"]] if p["a_index"] is not None else x[:R]
    out = own @ p["loop_w"].astype(np.float64)
    if graph:
        W = p["weight"].astype(np.float64).reshape(-1, D // S, S, S)
        dst = np.repeat(np.arange(R), np.diff(p["row_ptr"]))
        msg = np.einsum("ebi,ebij->ebj", x[p["e_src"]].reshape(-1, D // S, S), W[p["e_rel"]]).reshape(-1, D)
        agg = np.zeros((R, D))
        np.add.at(agg, dst, msg * p["norm"][dst, None].astype(np.float64))
        out += agg * p["norm"][:, None].astype(np.float64)
    if residual:
        out += own
    out += p["bias"].astype(np.float64)
    if act:
        out = np.maximum(out, 0.0)
    te = p["te"].astype(np.float64)[p["row_time"]]
    h = out + te if te_out else out
    ch = None
    if p["chain_n"]:
        ch = (out + te if te_chain else out) @ p["chain_w"].astype(np.float64) + p["chain_b"].astype(np.float64)
    return h, ch


def _run(p, act, residual, te_out, te_chain, packed, graph=True, lists=True):
    from temp_b200 import lib
    L = lib.load()
    dev = "cuda"
    t = {k: torch.from_numpy(v).to(dev) for k, v in p.items() if isinstance(v, np.ndarray)}
    for k, v in p["lists"].items():
        t["agg_" + k] = torch.from_numpy(v).to(dev)
    D, R, N = p["D"], p["R"], p["chain_n"]
    keep = []

    def pack(w):
        k, n = int(w.shape[0]), int(w.shape[1])
        nb = L.temp_packed_weights_bytes(k, n)
        assert nb > 0, (k, n)
        buf = torch.empty(nb, dtype=torch.uint8, device=dev)
        lib.check(L.temp_pack_weights(C.c_void_p(w.data_ptr()), k, n, C.c_void_p(buf.data_ptr()), C.c_void_p(lib.current_stream())), "pack")
        keep.append(buf)
        return buf.data_ptr()

    a = lib.RgcnLayerArgs()
    a.row0, a.row1, a.d = p["row0"], R, D
    agg = torch.full((R, D), float("nan"), device=dev)          # rows without in-edges must never be read
    if graph:
        a.row_ptr, a.e_src, a.e_rel, a.norm = t["row_ptr"].data_ptr(), t["e_src"].data_ptr(), t["e_rel"].data_ptr(), t["norm"].data_ptr()
        a.x, a.weight = t["x"].data_ptr(), t["weight"].data_ptr()
        a.n_bases, a.si, a.so = D // p["S"], p["S"], p["S"]
        a.agg_scratch = agg.data_ptr()
        if lists:
            a.agg_lists = 1
            a.agg_rows, a.n_agg_rows = t["agg_rows"].data_ptr(), int(t["agg_rows"].shape[0])
            a.agg_heavy, a.n_agg_heavy = t["agg_heavy"].data_ptr(), int(t["agg_heavy"].shape[0])
    a.residual, a.n_terms = int(residual), 1
    a.terms[0].a, a.terms[0].w = t["x"].data_ptr(), t["loop_w"].data_ptr()
    if p["a_index"] is not None:
        a.terms[0].a_index = t["a_index"].data_ptr()
    if packed:
        a.terms[0].w_packed = pack(t["loop_w"])
    a.h_bias = t["bias"].data_ptr()
    a.activation = lib.ACT_RELU if act else lib.ACT_NONE
    a.time_embed, a.row_time = t["te"].data_ptr(), t["row_time"].data_ptr()
    a.te_out, a.te_chain = int(te_out), int(te_chain)
    h = torch.full((R, D), -777.0, device=dev)
    a.h_out = h.data_ptr()
    ch = None
    if N:
        ch = torch.full((R, N + 4), -777.0, device=dev)
        a.chain_w, a.chain_b, a.chain_out = t["chain_w"].data_ptr(), t["chain_b"].data_ptr(), ch.data_ptr()
        a.chain_n, a.chain_ld = N, N + 4
        if packed:
            a.chain_w_packed = pack(t["chain_w"])
    a.inv_temperature = 0.1
    lib.check(L.temp_rgcn_layer_fwd(C.byref(a), C.c_void_p(lib.current_stream())), "temp_rgcn_layer_fwd")
    torch.cuda.synchronize()
    return h.cpu().numpy(), None if ch is None else ch.cpu().numpy()


def _err(got, want):
    return float(np.abs(got.astype(np.float64) - want).max() / max(np.abs(want).max(), 1e-30))


SHAPES = [  # D, S, R, row0, chain_n
    (200, 2, 333, 37, 600),      # BASELINE config 3: one GRU's input gates
    (200, 2, 200, 0, 1200),      # ... the Bi centre step: both cells' input gates (10 feature blocks, 3 ring slots reused)
    (200, 4, 130, 64, 0),        # no chain (layer 1)
    (32, 4, 100, 5, 96),         # one k-atom, quarter of a feature block
    (64, 1, 257, 0, 192),
    (160, 2, 70, 0, 480),
    (256, 4, 129, 1, 768),       # the widest operand (8 k-atoms)
    (128, 2, 300, 11, 384),      # d = 128 with 2x2 blocks (n_bases = 64)
    (100, 1, 90, 0, 300),        # d % 32 != 0 below one feature block
]


@pytest.mark.parametrize("shape", SHAPES, ids=["d%d_s%d_r%d_n%d" % (s[0], s[1], s[2], s[4]) for s in SHAPES])
def test_wide_layer_matches_fp64_and_the_simt_kernel(shape):
    D, S, R, row0, N = shape
    p = _problem(D, S, R, row0, N, seed=1000 + D + S)
    for act, te_out, te_chain in ((True, True, False), (False, False, True)):
        want_h, want_c = _reference(p, act, False, te_out, te_chain)
        res = {}
        for packed in (True, False):
            h, ch = _run(p, act, False, te_out, te_chain, packed)
            assert (h[:row0] == -777.0).all(), "rows below row0 were written"
            e = {"h": _err(h[row0:], want_h[row0:])}
            if N:
                assert (ch[:row0] == -777.0).all() and (ch[:, N:] == -777.0).all(), "chain output outside [row0, R) x [0, chain_n)"
                e["chain"] = _err(ch[row0:, :N], want_c[row0:])
            res[packed] = e
        _log({"shape": list(shape), "act": act, "te_out": te_out, "te_chain": te_chain, "tc": res[True], "simt": res[False]})
        for packed in (True, False):
            for k, v in res[packed].items():
                assert v < 1e-5, "%s path, %s: rel-to-scale error %.3e (%s)" % ("tcgen05" if packed else "SIMT", k, v, res)


def test_wide_layer_isolated_pass_with_indirection_and_without_work_lists():
    """forward_isolated (no aggregation, residual), the layer-1 indirection through a_index, and a launch without work lists."""
    p = _problem(200, 2, 150, 0, 600, seed=7, n_src=400)
    want_h, want_c = _reference(p, True, True, True, True, graph=False)
    h, ch = _run(p, True, True, True, True, packed=True, graph=False)
    assert _err(h, want_h) < 1e-5 and _err(ch[:, :600], want_c) < 1e-5, (_err(h, want_h), _err(ch[:, :600], want_c))
    want_h, want_c = _reference(p, False, False, False, False)
    for lists in (True, False):
        h, ch = _run(p, False, False, False, False, packed=True, lists=lists)
        assert _err(h, want_h) < 1e-5 and _err(ch[:, :600], want_c) < 1e-5, (lists, _err(h, want_h), _err(ch[:, :600], want_c))


def test_wide_gather_entry_is_bit_identical_to_the_simt_aggregation():
    """temp_rgcn_gather_fwd at d = 200 / 2x2 blocks: the same arithmetic in the same order as rgcn_layer_kernel's
    aggregation -- with zero self-loop weights, zero bias and no activation the SIMT layer's output IS its aggregate."""
    from temp_b200 import lib
    p = _problem(200, 2, 190, 0, 0, seed=11, heavy_at=1000, n_heavy=0)          # no heavy rows: one warp per row, edge order
    p["loop_w"] = np.zeros_like(p["loop_w"])
    p["bias"] = np.zeros_like(p["bias"])
    simt, _ = _run(p, False, False, False, False, packed=False)
    L = lib.load()
    t = {k: torch.from_numpy(v).cuda() for k, v in p.items() if isinstance(v, np.ndarray)}
    rows = torch.from_numpy(p["lists"]["rows"]).cuda()
    agg = torch.zeros(p["R"], 200, device="cuda")
    a = lib.RgcnLayerArgs()
    a.row0, a.row1, a.d = 0, p["R"], 200
    a.row_ptr, a.e_src, a.e_rel, a.norm = t["row_ptr"].data_ptr(), t["e_src"].data_ptr(), t["e_rel"].data_ptr(), t["norm"].data_ptr()
    a.x, a.weight, a.n_bases, a.si, a.so = t["x"].data_ptr(), t["weight"].data_ptr(), 100, 2, 2
    a.agg_scratch, a.agg_lists, a.agg_rows, a.n_agg_rows = agg.data_ptr(), 1, rows.data_ptr(), int(rows.shape[0])
    lib.check(L.temp_rgcn_gather_fwd(C.byref(a), C.c_void_p(lib.current_stream())), "temp_rgcn_gather_fwd")
    torch.cuda.synchronize()
    got = agg.cpu().numpy()
    _log({"gather_entry_bit_identical_to_simt": bool(np.array_equal(got, simt)), "err": _err(got, simt.astype(np.float64))})
    assert np.array_equal(got, simt)


def _gru_case(D, R, row0, seed, rec=True, accumulate=False, dt=True):
    g = np.random.default_rng(seed)
    f = lambda *s, scale=0.5: (g.standard_normal(s) * scale).astype(np.float32)
    n_prev = R + 17
    prev = g.integers(0, n_prev, size=R).astype(np.int32)
    prev[g.random(R) < 0.3] = -1
    return dict(D=D, R=R, row0=row0, gi=f(R, 3 * D + 8), state=f(n_prev, D), prev=prev if rec else None,
                dt=(g.integers(1, 5, size=R).astype(np.float32) if dt else None), whh_t=f(D, 3 * D, scale=1.0 / np.sqrt(D)),
                b_hh=f(3 * D), te=f(4, D), row_time=np.sort(g.integers(0, 4, size=R)).astype(np.int32),
                out0=f(R, D), accumulate=accumulate)


def _gru_reference(c):
    D, R = c["D"], c["R"]
    gi = c["gi"].astype(np.float64)[:, 4:4 + 3 * D]
    h0 = np.zeros((R, D))
    if c["prev"] is not None:
        ok = c["prev"] >= 0
        h0[ok] = c["state"].astype(np.float64)[c["prev"][ok]]
        if c["dt"] is not None:
            h0 *= np.exp(-c["dt"].astype(np.float64) * 0.1)[:, None]
    gh = h0 @ c["whh_t"].astype(np.float64) + c["b_hh"].astype(np.float64)
    sg = lambda x: 1.0 / (1.0 + np.exp(-x))
    r, z = sg(gi[:, :D] + gh[:, :D]), sg(gi[:, D:2 * D] + gh[:, D:2 * D])
    n = np.tanh(gi[:, 2 * D:] + r * gh[:, 2 * D:])
    hy = (1 - z) * n + z * h0 + c["te"].astype(np.float64)[c["row_time"]]
    return c["out0"].astype(np.float64) + hy if c["accumulate"] else hy


def _gru_run(c, packed):
    from temp_b200 import lib
    L = lib.load()
    D, R = c["D"], c["R"]
    t = {k: torch.from_numpy(v).cuda() for k, v in c.items() if isinstance(v, np.ndarray)}
    out = t["out0"].clone()
    a = lib.GruArgs()
    a.row0, a.row1, a.d = c["row0"], R, D
    a.gi, a.gi_ld, a.gi_off = t["gi"].data_ptr(), 3 * D + 8, 4
    if c["prev"] is not None:
        a.state, a.prev_row = t["state"].data_ptr(), t["prev"].data_ptr()
        if c["dt"] is not None:
            a.dt = t["dt"].data_ptr()
    a.inv_temperature = 0.1
    a.whh_t, a.b_hh = t["whh_t"].data_ptr(), t["b_hh"].data_ptr()
    if packed:
        nb = L.temp_packed_gru_bytes(D)
        assert nb > 0
        buf = torch.empty(nb, dtype=torch.uint8, device="cuda")
        lib.check(L.temp_pack_gru_weights(C.c_void_p(t["whh_t"].data_ptr()), D, C.c_void_p(buf.data_ptr()), C.c_void_p(lib.current_stream())), "pack gru")
        a.whh_packed = buf.data_ptr()
    a.cell_type = lib.CELL_TORCH_GRU
    a.time_embed, a.row_time = t["te"].data_ptr(), t["row_time"].data_ptr()
    a.accumulate, a.out, a.out_index_is_row = int(c["accumulate"]), out.data_ptr(), 1
    lib.check(L.temp_gru_fwd(C.byref(a), C.c_void_p(lib.current_stream())), "temp_gru_fwd")
    torch.cuda.synchronize()
    return out.cpu().numpy()


@pytest.mark.parametrize("D,R,row0", [(200, 333, 37), (200, 64, 0), (32, 100, 3), (160, 70, 0), (256, 129, 1), (100, 90, 0)])
def test_wide_gru_step_matches_fp64_and_the_simt_kernel(D, R, row0):
    """temp_gru_fwd at d != 128: gru_step_tcw_kernel (packed W_hh image given) and the fp32 SIMT gru_kernel against an fp64
    statement of the torch.nn.GRU step (models/RRGCN.py:77-89)."""
    for kw in (dict(), dict(accumulate=True, dt=False), dict(rec=False)):
        c = _gru_case(D, R, row0, seed=50 + D + R, **kw)
        want = _gru_reference(c)
        errs = {}
        for packed in (True, False):
            got = _gru_run(c, packed)
            assert np.array_equal(got[:row0], c["out0"][:row0]), "rows below row0 were written"
            errs["tc" if packed else "simt"] = _err(got[row0:], want[row0:])
        _log({"gru": [D, R, row0], "case": {k: str(v) for k, v in kw.items()}, **errs})
        assert errs["tc"] < 1e-5 and errs["simt"] < 1e-5, errs


@pytest.mark.parametrize("name", ["grrgcn_tiny_d200_nb100", "bigrrgcn_icews0515_real_nb100"])
def test_wide_scan_synchronisation_modes_agree(name):
    """gru_scan_tcw_kernel orders its steps by per-tile completion counters when the caller's zeroed words hold them
    (TempGruScanArgs.barrier_words) and by a grid-wide barrier otherwise; without a barrier word the scan runs as one
    gru_step_tcw_kernel launch per step.  Same arithmetic in the first two (bit-identical), fp32 rounding apart in the third."""
    from temp_b200 import lib
    from tests.helpers import CASE_BY_NAME, product_model
    case = CASE_BY_NAME[name]
    model = product_model(case)
    res = model.encode(case["t_list"])
    torch.cuda.synchronize()
    want = res.out.clone()
    scans = [o for o in res.program.ops if o.kind == lib.OP_GRU_SCAN]
    assert scans and all(o.u.scan.barrier_words > 2 for o in scans)
    for o in scans:
        o.u.scan.barrier_words = 2                     # only the two barrier words: grid-wide barrier between the steps
    res.program._arr = None
    res.out.zero_()
    res.program.run()
    torch.cuda.synchronize()
    assert torch.equal(res.out, want)
    for o in scans:
        o.u.scan.barrier = None                        # no barrier word at all: one launch per step
    res.program._arr = None
    res.out.zero_()
    res.program.run()
    torch.cuda.synchronize()
    assert float((res.out - want).abs().max()) <= 2e-6 * float(want.abs().max())
