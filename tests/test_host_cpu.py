"""CPU: host-side logic of the product (snapshots, planner, sampler, ABI surface) against the oracle
and the golden vectors.  No compute call is made without a GPU."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import temp_oracle as orc
from tests.golden.cases import CASES, SAMPLER_CASES
from tests.helpers import load_golden, oracle_graphs, oracle_model, product_store

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("dataset", ["tiny", "icews14_head"])
def test_snapshot_store_equals_reference_graph_construction(dataset):
    m, r, train, valid, test = oracle_graphs(dataset)
    store = product_store(dataset)
    assert (store.num_ents, store.num_rels) == (m, r)
    assert list(store.train.keys()) == list(train.keys())
    for split_p, split_o in ((store.train, train), (store.valid, valid), (store.test, test)):
        for t, g in split_o.items():
            s = split_p[t]
            assert np.array_equal(s.node_ids, g.ids)
            assert np.array_equal(s.src, g.src) and np.array_equal(s.dst, g.dst) and np.array_equal(s.rel, g.rel)
            assert np.array_equal(s.norm, g.norm)
            # CSR view: stable by destination => per-row edge order == edge-id order
            order = np.argsort(g.dst, kind="stable")
            assert np.array_equal(s.csr_src, g.src[order]) and np.array_equal(s.csr_rel, g.rel[order])
            assert np.array_equal(np.diff(s.row_ptr), np.bincount(g.dst, minlength=g.num_nodes))


def _plan_for(case):
    from temp_b200.planner import plan_static, plan_window
    store = product_store(case["dataset"])
    if case["module"] == "SRGCN":
        return plan_static(store.train, case["t_list"])
    return plan_window(store.train, case["t_list"], case["L"], bidirectional=case["module"].startswith("Bi"),
                       attention=case["module"].endswith("SARGCN"))


@pytest.mark.parametrize("case", [c for c in CASES if c["module"] in ("GRRGCN", "BiGRRGCN")],
                         ids=lambda c: c["name"])
def test_prev_row_encodes_the_dense_history(case):
    """prev_row >= 0 exactly where the oracle's dense history row is non-zero (history forgets)."""
    plan = _plan_for(case)
    model = oracle_model(case)
    with torch.no_grad():
        res = model.evaluate_embed(case["t_list"])
    assert plan.final_times == res["times"]
    hists = [res["hist"]] if "hist" in res else [res["hist_f"], res["hist_b"]]
    prevs = [plan.prev_a] if len(hists) == 1 else [plan.prev_a, plan.prev_b]
    for hist, prev in zip(hists, prevs):
        for i, inst in enumerate(plan.final.instances):
            nz = (hist[i][1][torch.from_numpy(inst.snapshot.node_ids)].abs().sum(-1) != 0).numpy()
            assert np.array_equal(prev[inst.row0:inst.row0 + inst.n] >= 0, nz)


@pytest.mark.parametrize("case", [c for c in CASES if c["module"] in ("GRRGCN", "BiGRRGCN", "RRGCN")],
                         ids=lambda c: c["name"])
@pytest.mark.parametrize("tile", [64, 5])
def test_chain_partitions_are_closed_and_cover_every_row(case, tile):
    """Every packed row belongs to exactly one chain partition, a partition has at most ``tile`` rows per segment
    and the recurrence never leaves it: prev_row of a row lies in a range of the same partition."""
    from temp_b200.planner import chain_partitions
    plan = _plan_for(case)
    parts = chain_partitions(plan, tile)
    assert parts.dtype == np.int32 and parts.shape[1:] == (len(plan.segments), 2)
    owner = np.full(plan.R, -1, dtype=np.int64)
    for p in range(parts.shape[0]):
        for g, seg in enumerate(plan.segments):
            lo, hi = parts[p, g]
            assert 0 <= hi - lo <= tile
            if hi > lo:
                assert seg.row0 <= lo and hi <= seg.row1
                assert (owner[lo:hi] == -1).all()
                owner[lo:hi] = p
    assert (owner >= 0).all()
    for prev in (plan.prev_a, plan.prev_b):
        if prev is None:
            continue
        has = prev >= 0
        assert np.array_equal(owner[prev[has]], owner[has])
    sizes = (parts[:, :, 1] - parts[:, :, 0]).sum(axis=1)
    assert (np.diff(sizes) <= 0).all()            # heaviest first


def test_plan_blob_roundtrip_and_alignment():
    plan = _plan_for(CASES[2])
    blob, lay, total = plan.to_blob()
    assert blob.nbytes == total
    for name, (off, nb) in lay.items():
        assert off % 256 == 0
        arr = getattr(plan, name)
        assert np.array_equal(blob[off:off + nb].view(arr.dtype).reshape(arr.shape), arr)
    assert plan.row_ptr[-1] == plan.E and plan.e_src.max() < plan.R
    assert plan.E == sum(i.snapshot.num_edges for s in plan.segments for i in s.instances)


def test_attention_slot_rows_match_oracle_mask():
    case = [c for c in CASES if c["name"] == "bisargcn_tiny_d128_last"][0]
    plan = _plan_for(case)
    with torch.no_grad():
        res = oracle_model(case).evaluate_embed(case["t_list"])
    mask = res["mask"]                                              # [2L-1, B, M]
    off = 0
    for i, inst in enumerate(plan.final.instances):
        active = (mask[:-1, i][:, torch.from_numpy(inst.snapshot.node_ids)] == 0).numpy().T    # [n, slots]
        assert np.array_equal(plan.slot_row[off:off + inst.n] >= 0, active)
        off += inst.n


@pytest.mark.parametrize("case", SAMPLER_CASES, ids=[c["name"] for c in SAMPLER_CASES])
def test_product_sampler_bit_exact_with_reference(case):
    from argparse import Namespace
    from temp_b200.sampler import CorruptTriples
    gold = load_golden(case["name"])
    store = product_store(case["dataset"])
    cor = CorruptTriples(Namespace(negative_rate=case["negative_rate"], num_pos_facts=case["num_pos_facts"]), store.train)
    np.random.seed(case["seed"])
    torch.manual_seed(case["seed"])
    for t in case["times"]:
        tri, nt, nh, lab = cor.single_graph_negative_sampling(t, store.train[t], store.num_ents)
        assert np.array_equal(tri.numpy(), gold["triples_%d" % t])
        assert np.array_equal(nt.numpy(), gold["neg_tail_%d" % t])
        assert np.array_equal(nh.numpy(), gold["neg_head_%d" % t])


def test_synthetic_generator_shape_statistics():
    from temp_b200.snapshot import SnapshotStore
    st = SnapshotStore.synthetic("icews14", num_times=12, scale=1, seed=1)
    e = np.mean([g.num_edges for g in st.train.values()])
    n = np.mean([g.num_nodes for g in st.train.values()])
    zero = np.mean([(g.norm == 0).mean() for g in st.train.values()])
    assert 120 < e < 280 and 140 < n < 320 and 0.3 < zero < 0.6
    st2 = SnapshotStore.synthetic("icews14", num_times=12, scale=1, seed=1)
    assert all(np.array_equal(st.train[t].src, st2.train[t].src) for t in st.train)
    gd = SnapshotStore.synthetic("gdelt", num_times=3, scale=1, seed=2)
    assert max(np.diff(g.row_ptr).max() for g in gd.train.values()) > 300


# ---- ABI surface -------------------------------------------------------------------------------
def _header_functions():
    text = open(os.path.join(ROOT, "include", "temp_b200.h")).read()
    return sorted(set(re.findall(r"^(?:int|int64_t|const char\*|TempPlan\*|void|const void\*)\s+(temp_\w+)\(", text, flags=re.M)))


def test_library_builds_loads_and_exports_every_declared_symbol():
    from temp_b200 import build, lib
    build.build()
    cdll = ctypes.CDLL(build.LIB_PATH)
    names = _header_functions()
    assert len(names) >= 10
    for name in names:
        assert hasattr(cdll, name), name
    assert sorted(lib.EXPORTS) == names
    assert lib.load().temp_abi_version() == lib.ABI_VERSION


def test_ctypes_structs_match_the_c_header(tmp_path):
    """sizeof/offsetof of every ABI struct as seen by gcc == the ctypes mirror."""
    from temp_b200 import lib
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "temp_b200.h"\nint main(){'
                   'printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(TempDenseTerm), sizeof(TempRgcnLayerArgs),'
                   'sizeof(TempGruArgs), sizeof(TempAttnArgs), sizeof(TempGatherArgs), sizeof(TempScatterArgs), sizeof(TempOp),'
                   'offsetof(TempRgcnLayerArgs, chain_ld), offsetof(TempGruArgs, out), offsetof(TempOp, u),'
                   'sizeof(TempGruScanArgs), offsetof(TempGruScanArgs, steps), sizeof(TempScoreLossArgs), sizeof(TempSnapshotView),'
                   'sizeof(TempPlanCounts), offsetof(TempRgcnLayerArgs, agg_lists), sizeof(TempRankArgs), offsetof(TempRankArgs, rank), sizeof(TempScoreLossBwdArgs), offsetof(TempScoreLossBwdArgs, grad_table));return 0;}')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    want = [ctypes.sizeof(lib.DenseTerm), ctypes.sizeof(lib.RgcnLayerArgs), ctypes.sizeof(lib.GruArgs),
            ctypes.sizeof(lib.AttnArgs), ctypes.sizeof(lib.GatherArgs), ctypes.sizeof(lib.ScatterArgs),
            ctypes.sizeof(lib.Op), lib.RgcnLayerArgs.chain_ld.offset, lib.GruArgs.out.offset, lib.Op.u.offset,
            ctypes.sizeof(lib.GruScanArgs), lib.GruScanArgs.steps.offset, ctypes.sizeof(lib.ScoreLossArgs),
            ctypes.sizeof(lib.SnapshotView), ctypes.sizeof(lib.PlanCounts), lib.RgcnLayerArgs.agg_lists.offset,
            ctypes.sizeof(lib.RankArgs), lib.RankArgs.rank.offset, ctypes.sizeof(lib.ScoreLossBwdArgs),
            lib.ScoreLossBwdArgs.grad_table.offset]
    assert got == want


@pytest.mark.parametrize("neg,num_ents,spread", [(500, 7128, True), (500, 7128, False), (64, 900, True), (5, 40, False)])
def test_native_negative_sampler_is_bit_identical_to_the_python_loop(neg, num_ents, spread):
    """temp_negative_sample continues NumPy's legacy MT19937 stream natively: the sample matrices AND the generator state
    afterwards equal the reference-style python loop (np.random.randint per round + np.isin) -- on a graph whose filter
    lists hit all three of numpy's in1d algorithms: short lists (per-value loop), long lists over a narrow id range
    (integer table), long lists over a wide id range (merge sort with assume_unique: candidates with a later twin in the
    round are dropped too)."""
    import torch
    from argparse import Namespace
    from temp_b200.sampler import CorruptTriples
    from temp_b200.snapshot import Snapshot
    rng = np.random.default_rng(neg + num_ents)
    n = min(300, num_ents)
    ids = np.sort(rng.choice(num_ents, size=n, replace=False)) if spread else np.arange(n) + (num_ents - n) // 2
    hub_tails = rng.choice(n, size=min(60, n - 1), replace=False)               # (h = 0, r = 0) has up to 60 true tails
    hub_heads = rng.choice(n, size=min(40, n - 1), replace=False)               # (r = 1, t = 1) has up to 40 true heads
    src = np.concatenate([np.zeros(hub_tails.size, dtype=np.int64), hub_heads, rng.integers(0, n, 80)])
    dst = np.concatenate([hub_tails, np.ones(hub_heads.size, dtype=np.int64), rng.integers(0, n, 80)])
    rel = np.concatenate([np.zeros(hub_tails.size, dtype=np.int64), np.ones(hub_heads.size, dtype=np.int64), rng.integers(0, 4, 80)])
    g = Snapshot(7, ids, src, dst, rel)
    for num_pos_facts in (3000, 50):                                            # 50 < E: torch.randperm subsampling first
        c = CorruptTriples(Namespace(negative_rate=neg, num_pos_facts=num_pos_facts), {7: g})
        for seed in (0, 1, 2):
            np.random.seed(seed); torch.manual_seed(seed); np.random.randn(seed + 1)
            a = c.single_graph_negative_sampling_python(7, g, num_ents)
            end_a = (np.random.randint(1 << 30, size=3).tolist(), np.random.randn(), torch.rand(2))
            np.random.seed(seed); torch.manual_seed(seed); np.random.randn(seed + 1)
            b = c.single_graph_negative_sampling(7, g, num_ents)
            end_b = (np.random.randint(1 << 30, size=3).tolist(), np.random.randn(), torch.rand(2))
            for x, y in zip(a, b):
                assert x.dtype == y.dtype and torch.equal(x, y)
            assert end_a[0] == end_b[0] and end_a[1] == end_b[1] and torch.equal(end_a[2], end_b[2])


def test_snapshot_store_binary_cache_round_trip(tmp_path):
    from temp_b200.snapshot import SnapshotStore
    from tests.helpers import product_store
    store = product_store("tiny")
    path = str(tmp_path / "tiny.npz")
    store.to_npz(path)
    back = SnapshotStore.from_npz(path)
    assert (back.num_ents, back.num_rels, back.name) == (store.num_ents, store.num_rels, store.name)
    for a, b in ((store.train, back.train), (store.valid, back.valid), (store.test, back.test)):
        assert list(a.keys()) == list(b.keys())
        for t in a:
            for field in ("node_ids", "src", "dst", "rel", "norm", "row_ptr", "csr_src", "csr_rel"):
                assert np.array_equal(getattr(a[t], field), getattr(b[t], field)), (t, field)


def test_reference_style_checkpoint_loads(tmp_path):
    """test.py:403-406: torch.load(path) -> load_state_dict(checkpoint['state_dict']) -> on_load_checkpoint(checkpoint)."""
    import torch
    from argparse import Namespace
    from tests.helpers import CASE_BY_NAME, product_model
    case = CASE_BY_NAME["bigrrgcn_tiny_d128_last"]
    src = product_model(case, device="cpu")
    ckpt = {"epoch": 3, "global_step": 17, "state_dict": {k: v.clone() + 0.25 for k, v in src.state_dict().items()},
            "hparams": Namespace(module="BiGRRGCN"), "optimizer_states": []}
    path = str(tmp_path / "ref.ckpt")
    torch.save(ckpt, path)
    dst = product_model(case, device="cpu")
    res = dst.load_reference_checkpoint(path)
    assert not res.missing_keys and not res.unexpected_keys
    for k, v in dst.state_dict().items():
        assert torch.equal(v, ckpt["state_dict"][k]), k
    dst.load_state_dict(ckpt["state_dict"])                      # the reference's own three calls also work
    dst.on_load_checkpoint(ckpt)


def test_every_package_module_imports():
    import importlib
    import pkgutil
    import temp_b200
    names = [m.name for m in pkgutil.iter_modules(temp_b200.__path__) if not m.name.startswith("libtemp")]   # the C-ABI .so
    assert {"evaluation", "autograd_path", "isolated", "sharding", "planner", "runtime", "models", "lib"} <= set(names)
    for name in names:
        importlib.import_module("temp_b200." + name)


def test_filter_lists_are_the_csr_form_of_the_dense_mask():
    """EvaluationFilter.filter_lists (the rank kernel's input) == the rows of the dense mask of utils/evaluation.py:82-99."""
    import torch
    from temp_b200.evaluation import EvaluationFilter
    from temp_b200.scores import complex_score
    from tests.helpers import product_store
    store = product_store("tiny")
    ev = EvaluationFilter(None, complex_score, store.train, store.valid, store.test)
    checked = 0
    for t, g in store.valid.items():
        if g.num_edges == 0:
            continue
        src, dst = g.edges()
        samples = torch.stack([src, g.edata["type_s"], dst]).transpose(0, 1)
        for mode in ("head", "tail"):
            ptr, flat = ev.filter_lists(samples, t, g, mode)
            mask = ev._mask(samples, store.num_ents, t, g, mode).numpy()
            assert ptr.dtype == np.int32 and flat.dtype == np.int32 and ptr.shape[0] == samples.shape[0] + 1
            for q in range(samples.shape[0]):
                assert np.array_equal(flat[ptr[q]:ptr[q + 1]], np.nonzero(mask[q])[0])
            checked += samples.shape[0]
    assert checked > 0


def test_encoder_refuses_to_run_without_cuda():
    """No CPU fallback: the product path must fail loudly."""
    if torch.cuda.is_available():
        pytest.skip("needs a CPU-only process")
    from tests.helpers import product_model
    model = product_model(CASES[2], device="cpu")
    with pytest.raises(RuntimeError, match="CUDA only"):
        model.encode(CASES[2]["t_list"])


def test_state_dict_keys_match_the_reference_shapes():
    from tests.helpers import oracle_config, product_args
    from temp_b200.models import build_module
    for case in CASES:
        cfg = oracle_config(case)
        store = product_store(case["dataset"])
        model = build_module(product_args(case), store.num_ents, store.num_rels, store.train)
        got = {k: tuple(v.shape) for k, v in model.state_dict().items()}
        assert got == orc.param_shapes(cfg), case["name"]


# ---- snapshot sharding: host logic + the two exchanges over gloo, world_size 2 ------------------
@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("name", ["grrgcn_icews_d128_L8", "bigrrgcn_icews_d128_L8", "grrgcn_tiny_d128_last"])
def test_shard_plan_cuts_instances_and_partitions(name, world):
    from tests.helpers import CASE_BY_NAME
    from temp_b200.sharding import make_shard_plan
    plan = _plan_for(CASE_BY_NAME[name])
    sp = make_shard_plan(plan, world)
    assert sp.row_bounds[0] == 0 and sp.row_bounds[-1] == plan.R and (np.diff(sp.row_bounds) >= 0).all()
    starts = {inst.row0 for seg in plan.segments for inst in seg.instances} | {plan.R}
    assert all(int(b) in starts for b in sp.row_bounds)                 # blocks are whole snapshot instances
    assert sp.parts.shape == plan.scan_parts.shape and sp.part_offset[-1] == plan.scan_parts.shape[0]
    key = lambda t: sorted(map(tuple, t.reshape(t.shape[0], -1).tolist()))
    assert key(sp.parts) == key(plan.scan_parts)                        # same partitions, rank-major order
    fin = plan.final
    got = np.sort(np.concatenate(sp.final_rows))
    assert np.array_equal(got, np.arange(fin.row0, fin.row1))           # every final row has exactly one owner


def _gloo_worker(rank, world, port, name, out_q):
    import torch.distributed as dist
    from tests.helpers import CASE_BY_NAME
    from temp_b200.sharding import exchange_blocks, exchange_rows, make_shard_plan
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        plan = _plan_for(CASE_BY_NAME[name])
        sp = make_shard_plan(plan, world)
        truth = torch.arange(plan.R * 6, dtype=torch.float32).reshape(plan.R, 6) * 0.5 + 1.0
        # exchange 1: every rank fills only its block of rows
        gi = torch.full((plan.R, 6), float("nan"))
        lo, hi = sp.rows_of(rank)
        gi[lo:hi] = truth[lo:hi]
        exchange_blocks(gi, sp.row_bounds)
        ok1 = bool(torch.equal(gi, truth))
        # exchange 2: every rank fills only the final rows of its chain partitions
        state = torch.full((plan.R, 6), -7.0)
        mine = torch.as_tensor(sp.final_rows[rank])
        state[mine] = truth[mine]
        for _ in range(2):                                              # second call goes through the cache
            exchange_rows(state, sp.final_rows, rank, None, sp.cache)
        fin = plan.final
        ok2 = bool(torch.equal(state[fin.row0:fin.row1], truth[fin.row0:fin.row1]))
        ok3 = bool((state[:fin.row0] == -7.0).all())                    # nothing else is touched
        out_q.put((rank, ok1, ok2, ok3))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name", ["grrgcn_icews_d128_L8", "bigrrgcn_icews_d128_L8"])
def test_shard_exchanges_over_gloo_world_size_2(name):
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, name, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True, True, True), (1, True, True, True)]


# ---- native planner (temp_b200/csrc/planner.cpp) == its python statement, array for array ---------------------------
def _plans_equal(a, b):
    for name in ("ent_id", "row_time", "norm", "row_ptr", "e_src", "e_src_ent", "e_rel", "prev_a", "dt_a", "prev_b", "dt_b",
                 "slot_row", "scan_parts", "agg_rows", "agg_heavy"):
        x, y = getattr(a, name), getattr(b, name)
        assert (x is None) == (y is None), name
        if x is not None:
            assert x.dtype == y.dtype and x.shape == y.shape and np.array_equal(x, y), name
    assert (a.R, a.E, a.n_slots, a.final_times, a.final_sizes) == (b.R, b.E, b.n_slots, b.final_times, b.final_sizes)
    assert len(a.segments) == len(b.segments)
    key = lambda i: None if i is None else (i.item, i.step, i.direction, i.time, i.row0, i.n, id(i.snapshot))
    for sa, sb in zip(a.segments, b.segments):
        assert (sa.kind, sa.step, sa.row0, sa.row1) == (sb.kind, sb.step, sb.row0, sb.row1)
        assert [key(i) for i in sa.instances] == [key(i) for i in sb.instances]
    assert [key(i) for i in a.last_hist_f] == [key(i) for i in b.last_hist_f]
    assert [key(i) for i in a.last_hist_b] == [key(i) for i in b.last_hist_b]
    assert [id(s) for s in a.final_snapshots] == [id(s) for s in b.final_snapshots]
    if a.n_slots:
        for da, db in zip(a.steps_f + list(getattr(a, "steps_b", [])), b.steps_f + list(getattr(b, "steps_b", []))):
            assert {k: key(v) for k, v in da.items()} == {k: key(v) for k, v in db.items()}


@pytest.mark.parametrize("case", [c for c in CASES if c["module"] != "SRGCN"], ids=lambda c: c["name"])
def test_native_planner_equals_python_planner(case):
    from temp_b200.planner import plan_window, plan_window_native
    store = product_store(case["dataset"])
    kw = dict(bidirectional=case["module"].startswith("Bi"), attention=case["module"].endswith("SARGCN"))
    _plans_equal(plan_window_native(store.train, case["t_list"], case["L"], **kw),
                 plan_window(store.train, case["t_list"], case["L"], **kw))


@pytest.mark.parametrize("shape,L,B,bi,att", [("icews14", 8, 8, False, False), ("icews05-15", 8, 8, True, False),
                                               ("icews14", 8, 8, True, True), ("gdelt", 15, 2, False, False),
                                               ("icews14", 3, 5, False, True)])
def test_native_planner_equals_python_planner_on_synthetic_shapes(shape, L, B, bi, att):
    from temp_b200.planner import plan_window, plan_window_native
    from temp_b200.snapshot import SnapshotStore
    store = SnapshotStore.synthetic(shape, num_times=2 * L + B + 3, scale=1, seed=99)
    rng = np.random.default_rng(5)
    hi = len(store.times) - (L if bi else 1)
    for _ in range(3):
        t_list = [int(store.times[i]) for i in rng.choice(np.arange(0, hi + 1), size=B, replace=False)]
        _plans_equal(plan_window_native(store.train, t_list, L, bidirectional=bi, attention=att),
                     plan_window(store.train, t_list, L, bidirectional=bi, attention=att))


def test_planners_agree_on_the_chosen_scan_tile():
    """scan_tile < 0 = "48 rows per partition step, 64 when that leaves more than two rounds of the scan kernel's pipelines":
    both planners apply the rule alike -- a small batch keeps 48, a x5 batch is cut at 64 -- and report the tile they used."""
    from temp_b200.planner import AUTO_TILE_PARTS, chain_partitions, plan_window, plan_window_native
    from temp_b200.snapshot import SnapshotStore
    for scale, want_tile in ((1, 48), (5, 64)):
        store = SnapshotStore.synthetic("icews14", num_times=20, scale=scale, seed=7)
        t_list = [int(t) for t in store.times[8:16]]
        a = plan_window_native(store.train, t_list, 8, scan_tile=-48)
        b = plan_window(store.train, t_list, 8, scan_tile=-48)
        assert a.scan_tile == b.scan_tile == want_tile
        _plans_equal(a, b)
        n48 = chain_partitions(b, 48).shape[0]
        assert (n48 > AUTO_TILE_PARTS) == (want_tile == 64)
        rows = b.scan_parts[:, :, 1] - b.scan_parts[:, :, 0]
        assert rows.max() <= want_tile and (want_tile == 48 or rows.max() > 48)


def test_model_shell_answers_the_reference_hook_names_and_loaders():
    """The hook / loader surface PL and test.py call on a TKG_Module (models/TKG_Module.py:43-200) exists on the shells;
    the loaders yield batches of ``batch_size`` target timestamps."""
    from tests.helpers import CASE_BY_NAME, product_args
    from temp_b200.models import build_module
    case = CASE_BY_NAME["grrgcn_tiny_d128_last"]
    store = product_store(case["dataset"])
    args = product_args(case)
    args.batch_size = 4
    model = build_module(args, store.num_ents, store.num_rels, store.train, store.valid, store.test)
    for name in ("forward", "evaluate", "evaluate_embed", "train_embed", "get_all_embeds_Gt", "calc_metrics", "training_step",
                 "validation_step", "test_step", "validation_end", "test_end", "configure_optimizers", "train_dataloader",
                 "val_dataloader", "test_dataloader", "train_link_prediction", "link_classification_loss",
                 "get_batch_graph_list", "get_metrics"):
        assert callable(getattr(model, name)), name
    batches = list(model.val_dataloader())
    assert sum(int(b.numel()) for b in batches) == len(store.times) and all(int(b.numel()) <= 4 for b in batches)
    assert sorted(int(t) for b in batches for t in b) == sorted(int(t) for t in store.times)
    out = model.test_end([{"ranks": torch.tensor([1, 2, 10]), "test_loss": 0.5, "batch_time": batches[0]}])
    assert abs(out["mrr"] - (1 + 0.5 + 0.1) / 3) < 1e-6 and out["hit_1"] == pytest.approx(1 / 3)


def test_tensor_core_image_sizes_cover_the_wide_family_without_a_gpu():
    """temp_packed_weights_bytes / temp_packed_gru_bytes (no CUDA call): 32 KB chunks of 128 features x 32 k (hi + lo image);
    d = 128 keeps the round-1 image, every other d % 4 == 0 up to 256 pads k to whole k-atoms and n to whole 128-feature
    blocks (tc_wide.cu); shapes without a tensor-core path answer with a negative size; a scan made of such steps counts one
    launch (cooperative, d <= 224) and TempGruScanArgs carries the counters' word count."""
    import ctypes as C
    from temp_b200 import lib
    L = lib.load()
    chunk = 2 * 128 * 128
    assert L.temp_packed_weights_bytes(128, 384) == 3 * 4 * chunk
    assert L.temp_packed_weights_bytes(200, 200) == 2 * 7 * chunk
    assert L.temp_packed_weights_bytes(200, 1200) == 10 * 7 * chunk
    assert L.temp_packed_weights_bytes(32, 96) == 1 * 1 * chunk
    assert L.temp_packed_weights_bytes(256, 768) == 6 * 8 * chunk
    assert L.temp_packed_weights_bytes(130, 130) < 0 and L.temp_packed_weights_bytes(260, 260) < 0
    assert L.temp_packed_gru_bytes(128) == 4 * 4 * chunk
    assert L.temp_packed_gru_bytes(200) == 7 * 7 * chunk
    assert L.temp_packed_gru_bytes(258) < 0
    sc = lib.GruScanArgs()
    sc.n_steps, sc.barrier, sc.barrier_words = 3, 0x1000, 2 + 3 * 8
    for i in range(3):
        g = sc.steps[i]
        g.row0, g.row1, g.d = 64 * i, 64 * i + 50, 200
        g.gi, g.b_hh, g.out, g.whh_packed, g.state = 0x2000, 0x3000, 0x4000, 0x5000, 0x4000
        g.prev_row = 0x6000 if i else None
        g.cell_type = lib.CELL_TORCH_GRU
    op = lib.Op()
    op.kind = lib.OP_GRU_SCAN
    op.u.scan = sc
    arr = (lib.Op * 1)(op)
    assert L.temp_program_kernel_count(arr, 1) == 1
    arr[0].u.scan.steps[0].d = arr[0].u.scan.steps[1].d = arr[0].u.scan.steps[2].d = 256      # W_hh slice beyond tensor memory
    assert L.temp_program_kernel_count(arr, 1) == 3


def test_fused_scan_ops_carry_the_barrier_words():
    """Program.fuse_gru_scans: consecutive GRU steps become one scan op that carries the barrier pointer and the number of
    zeroed words behind it (TempGruScanArgs.barrier_words: what lets gru_scan_tcw_kernel use per-tile completion counters);
    with split_cells a run also ends where the recurrent cell changes."""
    from temp_b200 import lib
    prog = lib.Program()
    for i in range(5):
        g = lib.GruArgs()
        g.row0, g.row1, g.d, g.b_hh = 10 * i, 10 * i + 10, 200, 0x1000 if i < 3 else 0x2000
        prog.add(lib.OP_GRU, g)
    prog.fuse_gru_scans(0xabc0, barrier_words=2 + 16 * 1024)
    assert [o.kind for o in prog.ops] == [lib.OP_GRU_SCAN]
    sc = prog.ops[0].u.scan
    assert sc.n_steps == 5 and sc.barrier == 0xabc0 and sc.barrier_words == 2 + 16 * 1024
    assert [sc.steps[i].row0 for i in range(5)] == [0, 10, 20, 30, 40]
    prog2 = lib.Program()
    for i in range(5):
        g = lib.GruArgs()
        g.row0, g.row1, g.d, g.b_hh = 10 * i, 10 * i + 10, 128, 0x1000 if i < 3 else 0x2000
        prog2.add(lib.OP_GRU, g)
    prog2.fuse_gru_scans(0xabc0, split_cells=True)
    assert [o.kind for o in prog2.ops] == [lib.OP_GRU_SCAN, lib.OP_GRU_SCAN]
    assert [o.u.scan.n_steps for o in prog2.ops] == [3, 2] and prog2.ops[0].u.scan.barrier_words == 2
