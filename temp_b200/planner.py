"""Window planner: turns a batch of target timestamps into ONE packed, device-ready work list.

The reference rebuilds a batched DGL graph on the host and copies it to the GPU at every time step
(models/DynamicRGCN.py:76-110, utils/utils.py:9-11) and carries the recurrent state in a dense
``(B, 2, num_ents, D)`` tensor that is re-zeroed every step (DynamicRGCN.py:47-54).  Here the whole
window batch is laid out once as "packed rows" -- the nodes of every snapshot instance back to
back, step-major -- with

  * one CSR-by-destination over all packed rows (edge order inside a row = edge-id order),
  * ``prev_row``: for every row the packed row of the SAME entity in the SAME batch item at the
    previous step, or -1.  This is exactly what the dense history encodes, including the
    "history forgets" quirk: an entity absent from step k-1 has zero state at k,
  * ``slot_row`` (attention models): the packed history row of the entity per time slot, or -1.

Window construction follows models/TKG_Module.py:232-250 (forward: items sorted descending, the last
``seq_len`` timestamps <= t, None-padded at the front) and models/BiDynamicRGCN.py:17-49 (backward:
items sorted ascending, timestamps t .. t+L-1 reversed so that t comes last).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np

from .snapshot import Snapshot

__all__ = ["Instance", "Segment", "WindowPlan", "plan_window", "plan_static"]

_ALIGN = 256
# in-degree above which the aggregation kernel gives a destination row a whole thread block (0 = chosen per batch by its
# mean in-degree, see agg_heavy_degree; TEMP_AGG_HEAVY pins it for experiments)
AGG_HEAVY_DEGREE = int(__import__('os').environ.get('TEMP_AGG_HEAVY', '0'))


@dataclass
class Instance:
    item: int          # batch position in forward (descending-time) order
    step: int
    direction: str     # 'f' | 'b' | 'c' (centre / final step)
    time: int
    row0: int
    n: int
    snapshot: Snapshot


@dataclass
class Segment:
    kind: str          # 'hist_f' | 'hist_b' | 'final'
    step: int
    row0: int
    row1: int
    instances: List[Instance] = field(default_factory=list)


def agg_heavy_degree(n_edges: int, n_rows_with_edges: int) -> int:
    """In-degree above which the aggregation kernel gives a destination row a whole thread block instead of a warp: four
    times the batch's mean in-degree over the rows that have in-edges, within [8, 64] (csrc/planner.cpp computes the same).
    Measured on a B200 (tools/probe_config.py): sparse snapshots (ICEWS14, mean 1.6) want 8 -- a lone warp walking a 30-edge
    row is the tail of a 5 us launch (forward 86 -> 92 us at 32); dense ones (GDELT, mean 16.5) want 64 -- at 8 nearly every
    row takes a block of 8 warps for two edges each (forward 160 -> 139 us at 64, 154 us at 128)."""
    if AGG_HEAVY_DEGREE > 0:
        return AGG_HEAVY_DEGREE
    if n_rows_with_edges <= 0:
        return 8
    return int(min(64, max(8, (4 * n_edges) // n_rows_with_edges)))


class WindowPlan(object):
    """Packed arrays (numpy, int32 / float32) + segment table.  ``to_blob`` concatenates them into
    one byte buffer so that the host->device traffic of a forward is a single copy."""

    ARRAYS = ("ent_id", "row_time", "norm", "row_ptr", "e_src", "e_src_ent", "e_rel", "prev_a", "dt_a",
              "prev_b", "dt_b", "slot_row", "scan_parts", "agg_rows", "agg_heavy")

    def __init__(self):
        self.segments: List[Segment] = []
        self.seq_len = 0
        self.batch = 0
        self.bidirectional = False
        self.R = 0
        self.E = 0
        self.n_slots = 0
        self.scan_tile = 0            # row bound of one chain-partition step in scan_parts (0: no table)
        self.final_times: List[int] = []
        self.final_sizes: List[int] = []
        self.final_snapshots: List[Snapshot] = []
        # per item: Instance of the last history step (step L-2), forward / backward, or None
        self.last_hist_f: List[Optional[Instance]] = []
        self.last_hist_b: List[Optional[Instance]] = []
        for name in self.ARRAYS:
            setattr(self, name, None)

    @property
    def final(self) -> Segment:
        return self.segments[-1]

    @property
    def hist_rows(self) -> int:
        return self.final.row0

    def edges_processed(self) -> int:
        return int(self.E)

    def blob_layout(self):
        off, lay = 0, {}
        for name in self.ARRAYS:
            arr = getattr(self, name)
            if arr is None:
                continue
            lay[name] = (off, arr.nbytes)
            off = (off + arr.nbytes + _ALIGN - 1) // _ALIGN * _ALIGN
        return lay, off

    def to_blob(self, out: Optional[np.ndarray] = None):
        lay, total = self.blob_layout()
        if out is None:
            out = np.zeros(total, dtype=np.uint8)
        assert out.nbytes >= total
        for name, (off, nb) in lay.items():
            out[off:off + nb] = getattr(self, name).view(np.uint8).reshape(-1)
        return out, lay, total


def _window_times(t_list: Sequence[int], seq_len: int, times: List[int], backward: bool):
    pos = {t: i for i, t in enumerate(times)}
    order = sorted((int(t) for t in t_list), reverse=not backward)
    rows = []
    for tim in order:
        p = pos[tim]
        if backward:
            seq = list(reversed(times[p:p + seq_len]))
        else:
            seq = times[max(0, p + 1 - seq_len):p + 1]
        rows.append([None] * (seq_len - len(seq)) + list(seq))
    return rows  # rows[item][step]


class _Packer(object):
    def __init__(self):
        self.ent, self.rtime, self.norm, self.deg = [], [], [], []
        self.esrc, self.esrc_ent, self.erel = [], [], []
        self.prev_a, self.dt_a, self.prev_b, self.dt_b = [], [], [], []
        self.R = 0
        self.E = 0

    def add(self, snap: Snapshot) -> int:
        row0 = self.R
        n = snap.num_nodes
        ids32, rtime, deg, esrc_ent, _ = snap.packed_parts()
        self.ent.append(ids32)
        self.rtime.append(rtime)
        self.norm.append(snap.norm)
        self.deg.append(deg)
        self.esrc.append(snap.csr_src + np.int32(row0))
        self.esrc_ent.append(esrc_ent)
        self.erel.append(snap.csr_rel)
        self.R += n
        self.E += snap.num_edges
        return row0

    def add_prev(self, n: int, prev: Optional[np.ndarray], dt_val: float, which: str = "a"):
        pv = prev if prev is not None else np.full(n, -1, dtype=np.int32)
        dt = np.full(n, dt_val, dtype=np.float32)
        (self.prev_a if which == "a" else self.prev_b).append(pv.astype(np.int32))
        (self.dt_a if which == "a" else self.dt_b).append(dt)

    def finish(self, plan: WindowPlan):
        cat = lambda xs, dt: np.ascontiguousarray(np.concatenate(xs) if xs else np.zeros(0), dtype=dt)
        plan.R, plan.E = self.R, self.E
        plan.ent_id = cat(self.ent, np.int32)
        plan.row_time = cat(self.rtime, np.int32)
        plan.norm = cat(self.norm, np.float32)
        rp = np.zeros(self.R + 1, dtype=np.int64)
        if self.deg:
            np.cumsum(np.concatenate(self.deg), out=rp[1:])
        plan.row_ptr = rp.astype(np.int32)
        plan.e_src = cat(self.esrc, np.int32)
        plan.e_src_ent = cat(self.esrc_ent, np.int32)
        plan.e_rel = cat(self.erel, np.int32)
        # work list of the aggregation kernel: (packed row, first edge, end edge) of every row WITH in-edges
        # (a warp each), split by in-degree: rows above AGG_HEAVY_DEGREE get a whole thread block
        deg = rp[1:] - rp[:-1]
        heavy = agg_heavy_degree(int(rp[-1]), int(np.count_nonzero(deg)))
        for name, nz in (("agg_rows", np.nonzero((deg > 0) & (deg <= heavy))[0]),
                         ("agg_heavy", np.nonzero(deg > heavy)[0])):
            setattr(plan, name, np.ascontiguousarray(np.stack([nz, rp[nz], rp[nz + 1]], axis=1), dtype=np.int32)
                    if nz.size else np.zeros((0, 3), dtype=np.int32))
            setattr(plan, name + "_ids", nz.astype(np.int64))
        plan.prev_a = cat(self.prev_a, np.int32)
        plan.dt_a = cat(self.dt_a, np.float32)
        if self.prev_b:
            plan.prev_b = cat(self.prev_b, np.int32)
            plan.dt_b = cat(self.dt_b, np.float32)


def _match_prev(cur: Snapshot, prev_inst: Optional[Instance]) -> Optional[np.ndarray]:
    """Packed row of each node of ``cur`` in the previous instance of the same item, -1 if absent."""
    if prev_inst is None:
        return None
    pid = prev_inst.snapshot.node_ids
    pos = np.searchsorted(pid, cur.node_ids)
    pos_c = np.minimum(pos, pid.shape[0] - 1)
    hit = pid[pos_c] == cur.node_ids
    return np.where(hit, pos_c + prev_inst.row0, -1).astype(np.int32)


def plan_window(graph_dict: Dict[int, Snapshot], t_list: Sequence[int], seq_len: int, bidirectional: bool = False,
                attention: bool = False, transform=None, scan_tile: Optional[int] = None) -> WindowPlan:
    """Plan the forward of (Bi)DynamicRGCN / (Bi)SelfAttentionRGCN for the target timestamps ``t_list``.

    Row order: forward history steps 0..L-2, then (Bi) backward history steps 0..L-2, then the final
    (centre) step whose graphs are the targets in descending-time order.

    ``transform(kind, snapshot) -> snapshot`` (kind 'hist' | 'final') substitutes the graph an instance is computed
    on -- the training-mode edge sub-sampling of models/DynamicRGCN.py:76-94 -- and is called in the reference's
    order: history steps outer, batch items inner, then the final step's items.  ``plan.final_snapshots`` keeps the
    full graphs (the negative sampler works on them, DynamicRGCN.py:185-187)."""
    times = list(graph_dict.keys())
    L, B = int(seq_len), len(t_list)
    plan = WindowPlan()
    plan.seq_len, plan.batch, plan.bidirectional = L, B, bidirectional
    pk = _Packer()
    fwd = _window_times(t_list, L, times, backward=False)            # [item][step], items descending
    bwd = _window_times(t_list, L, times, backward=True) if bidirectional else None   # items ascending

    def history(rows, kind, flip):
        last: List[Optional[Instance]] = [None] * B
        by_step: List[Dict[int, Instance]] = []
        for k in range(L - 1):
            seg = Segment(kind, k, pk.R, pk.R)
            cur: Dict[int, Instance] = {}
            for j in range(B):
                tim = rows[j][k]
                if tim is None:
                    continue
                snap = graph_dict[tim]
                if transform is not None:
                    snap = transform("hist", snap)
                item = (B - 1 - j) if flip else j                     # BiDynamicRGCN.py:97-99 (flip)
                prev = _match_prev(snap, last[j])
                inst = Instance(item, k, kind[-1], tim, pk.add(snap), snap.num_nodes, snap)
                # dt = cur_t - start_time; it only multiplies a non-zero state, i.e. rows whose
                # entity was active at step k-1, where it equals 1
                pk.add_prev(snap.num_nodes, prev, 1.0 if prev is not None else float(k))
                if bidirectional:
                    pk.add_prev(snap.num_nodes, None, float(k), "b")
                cur[j] = inst
                seg.instances.append(inst)
            seg.row1 = pk.R
            for j, inst in cur.items():
                last[j] = inst
            by_step.append(cur)
            if seg.row1 > seg.row0:
                plan.segments.append(seg)
        return last, by_step

    last_f, steps_f = history(fwd, "hist_f", flip=False)
    last_b, steps_b = (history(bwd, "hist_b", flip=True) if bidirectional else ([None] * B, []))

    seg = Segment("final", L - 1, pk.R, pk.R)
    slot_rows = []
    for i in range(B):
        tim = fwd[i][L - 1]
        full_snap = graph_dict[tim]
        snap = transform("final", full_snap) if transform is not None else full_snap
        inst = Instance(i, L - 1, "c", tim, pk.add(snap), snap.num_nodes, snap)
        pf = _match_prev(snap, last_f[i])
        pk.add_prev(snap.num_nodes, pf, 1.0 if pf is not None else float(L - 1))
        if bidirectional:
            pb = _match_prev(snap, last_b[B - 1 - i])
            pk.add_prev(snap.num_nodes, pb, 1.0 if pb is not None else float(L - 1), "b")
        seg.instances.append(inst)
        plan.final_times.append(tim)
        plan.final_sizes.append(snap.num_nodes)
        plan.final_snapshots.append(full_snap)
        if attention:
            cols = []
            for k in range(L - 1):
                m = _match_prev(snap, steps_f[k].get(i))
                cols.append(m if m is not None else np.full(snap.num_nodes, -1, dtype=np.int32))
            if bidirectional:
                for k in range(L - 1):
                    m = _match_prev(snap, steps_b[k].get(B - 1 - i))
                    cols.append(m if m is not None else np.full(snap.num_nodes, -1, dtype=np.int32))
            slot_rows.append(np.stack(cols, axis=1) if cols else np.zeros((snap.num_nodes, 0), dtype=np.int32))
    seg.row1 = pk.R
    plan.segments.append(seg)
    plan.last_hist_f = list(last_f)
    plan.last_hist_b = [last_b[B - 1 - i] for i in range(B)] if bidirectional else [None] * B
    pk.finish(plan)
    if not attention:
        plan.scan_tile, plan.scan_parts = partition_scan(plan, int(scan_tile or SCAN_TILE))
    if attention:
        plan.n_slots = (L - 1) * (2 if bidirectional else 1)
        plan.slot_row = np.ascontiguousarray(np.concatenate(slot_rows, axis=0), dtype=np.int32)
        plan.steps_f, plan.steps_b = steps_f, steps_b
    return plan


SCAN_TILE = 96      # rows per chain partition and step of gru_scan_tc_kernel (round-1 scan; the default when no model asks for 48)
SCAN_TILE_TM = 48   # ... of gru_scan_tm_kernel (one recurrent cell, W_hh in tensor memory; temp_b200/csrc/tc_scan2.cu)


AUTO_TILE_PARTS = 132     # 2 rounds of the scan kernel's pipelines (2 x 33 clusters x 2)


def partition_scan(plan: WindowPlan, tile: int):
    """-> (tile used, partition table).  ``tile`` < 0: |tile| rows per partition step, widened to 64 when that leaves more
    than AUTO_TILE_PARTS partitions -- the scan is then throughput bound and 25 % fewer, fuller tile steps win (x16 bench
    shape: 367 -> 334 us), while in the latency regime the narrower tile is faster (x1: 45 against 51 us).  Same rule as
    csrc/planner.cpp."""
    auto = tile < 0
    tile = abs(tile)
    parts = chain_partitions(plan, tile)
    if auto and tile < 64 and parts.shape[0] > AUTO_TILE_PARTS:
        tile = 64
        parts = chain_partitions(plan, tile)
    return tile, parts


def chain_partitions(plan: WindowPlan, tile: int = SCAN_TILE) -> np.ndarray:
    """Cuts every batch item into entity-id ranges such that no step of a range has more than ``tile`` rows.

    The recurrence links a row only to the row of the SAME entity in the SAME batch item at the previous step
    (``prev_row``; reference models/DynamicRGCN.py:35-54), and the rows of a snapshot instance are sorted by entity
    id (utils/dataset.py:168), so an entity-id range owns one contiguous packed-row range per segment and no
    dependency leaves it.  Returns int32 ``[P, len(plan.segments), 2]`` row ranges ``[lo, hi)`` (lo == hi: the
    partition has no rows in that segment), partitions with the most rows first."""
    n_seg = len(plan.segments)
    per_item: Dict[int, list] = {}
    for g, seg in enumerate(plan.segments):
        for inst in seg.instances:
            per_item.setdefault(inst.item, []).append((g, inst))
    parts = []
    big = np.iinfo(np.int64).max
    for item in sorted(per_item):
        insts = per_item[item]
        K = len(insts)
        sizes = np.array([inst.n for _, inst in insts], dtype=np.int64)
        segs = np.array([g for g, _ in insts], dtype=np.int64)
        row0 = np.array([inst.row0 for _, inst in insts], dtype=np.int64)
        ids = np.full((K, int(sizes.max()) + 1), big, dtype=np.int64)        # entity ids per instance, padded with +inf
        for k, (_, inst) in enumerate(insts):
            ids[k, :inst.n] = inst.snapshot.node_ids
        pos = np.zeros(K, dtype=np.int64)
        rows_k = np.arange(K)
        while (pos < sizes).any():
            # the first entity id that would make some instance exceed `tile` rows; everything below it joins the partition
            cut = ids[rows_k, np.minimum(pos + tile, sizes)].min()
            new = sizes if cut == big else (ids < cut).sum(axis=1)
            row = np.zeros((n_seg, 2), dtype=np.int32)
            row[segs, 0] = row0 + pos
            row[segs, 1] = row0 + new
            parts.append(row)
            pos = new
    if not parts:
        return np.zeros((0, n_seg, 2), dtype=np.int32)
    out = np.stack(parts)
    order = np.argsort(-(out[:, :, 1] - out[:, :, 0]).sum(axis=1), kind="stable")
    return np.ascontiguousarray(out[order])


# ---------------------------------------------------------------------------------------------------------------------
# native planner (temp_b200/csrc/planner.cpp): same algorithm, same arrays, ~20x faster than the numpy statement above
# ---------------------------------------------------------------------------------------------------------------------
_VIEW_CACHE: Dict[int, tuple] = {}


def _snapshot_views(graph_dict: Dict[int, Snapshot]):
    """ctypes table of TempSnapshotView for every snapshot of ``graph_dict`` (cached per dict object)."""
    import ctypes as C
    from . import lib
    key = id(graph_dict)
    hit = _VIEW_CACHE.get(key)
    if hit is not None and hit[0] is graph_dict and hit[1] == len(graph_dict):
        return hit[2], hit[3]
    times = list(graph_dict.keys())
    views = (lib.SnapshotView * len(times))()
    keep = []
    for i, t in enumerate(times):
        s = graph_dict[t]
        ids32 = s.packed_parts()[0]
        keep.append(ids32)
        v = views[i]
        v.time, v.n_nodes, v.n_edges = int(s.time), s.num_nodes, s.num_edges
        v.node_ids, v.row_ptr = ids32.ctypes.data, s.row_ptr.ctypes.data
        v.csr_src, v.csr_rel, v.norm = s.csr_src.ctypes.data, s.csr_rel.ctypes.data, s.norm.ctypes.data
    index = {t: i for i, t in enumerate(times)}
    _VIEW_CACHE[key] = (graph_dict, len(times), (views, keep, times), index)
    return (views, keep, times), index


_PLAN_DTYPES = {"norm": np.float32, "dt_a": np.float32, "dt_b": np.float32}
_KINDS = ("hist_f", "hist_b", "final")
_DIRS = ("f", "b", "c")


class NativeWindowPlan(WindowPlan):
    """A WindowPlan whose arrays live in the native planner's memory: the device blob is written straight from C
    (``to_blob``), numpy views are materialised only when somebody asks for an array by name (tests, the dense-history
    API, the autograd fallback)."""

    _LAZY = {name: i for i, name in enumerate(("ent_id", "row_time", "norm", "row_ptr", "e_src", "e_src_ent", "e_rel", "prev_a",
                                               "dt_a", "prev_b", "dt_b", "slot_row", "scan_parts", "agg_rows", "agg_heavy"))}

    def __init__(self, handle, lib_mod):
        self.__dict__["_handle"] = handle
        self.__dict__["_lib"] = lib_mod
        WindowPlan.__init__(self)
        for name in self.ARRAYS:               # WindowPlan.__init__ set them to None: make them lazy again
            self.__dict__.pop(name, None)
        self._absent = set()

    def __del__(self):
        h = self.__dict__.get("_handle")
        if h:
            try:
                self._lib.load().temp_plan_destroy(h)
            except Exception:
                pass
            self.__dict__["_handle"] = None

    def _fetch(self, which, dtype):
        import ctypes as C
        nb = C.c_int64()
        ptr = self._lib.load().temp_plan_array(self._handle, which, C.byref(nb))
        n = nb.value // 4
        if not n:
            return np.zeros(0, dtype=dtype)
        ct = C.c_float if dtype == np.float32 else C.c_int32
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), shape=(n,)).copy()

    def __setattr__(self, name, value):        # an array replaced from python (e.g. the rank-major partition table of a
        if name in type(self)._LAZY and "_absent" in self.__dict__:   # sharded forward) overrides the native copy
            self.__dict__.setdefault("_dirty", set()).add(name)
        self.__dict__[name] = value

    def __getattr__(self, name):               # only called when the attribute is not in __dict__
        lazy = type(self)._LAZY
        if name in lazy:
            if name in self.__dict__.get("_absent", ()):
                val = None
            else:
                val = self._fetch(lazy[name], _PLAN_DTYPES.get(name, np.int32))
                if name in ("agg_rows", "agg_heavy"):
                    val = val.reshape(-1, 3)
                elif name == "slot_row":
                    val = val.reshape(self.final.row1 - self.final.row0, self.n_slots)
                elif name == "scan_parts":
                    val = val.reshape(self._n_parts, len(self.segments), 2)
            self.__dict__[name] = val
            return val
        if name in ("agg_rows_ids", "agg_heavy_ids"):
            val = getattr(self, name[:-4])[:, 0].astype(np.int64)
            self.__dict__[name] = val
            return val
        raise AttributeError(name)

    def blob_layout(self):
        import ctypes as C
        n = len(self._LAZY)
        offs, sizes = (C.c_int64 * n)(), (C.c_int64 * n)()
        total = self._lib.load().temp_plan_blob_layout(self._handle, offs, sizes, _ALIGN)
        if total < 0:
            raise RuntimeError("temp_b200: temp_plan_blob_layout failed")
        lay = {name: (int(offs[i]), int(sizes[i])) for name, i in self._LAZY.items() if offs[i] >= 0}
        return lay, int(total)

    def to_blob(self, out: Optional[np.ndarray] = None):
        lay, total = self.blob_layout()
        if out is None:
            out = np.zeros(total, dtype=np.uint8)
        assert out.nbytes >= total and out.flags["C_CONTIGUOUS"]
        self._lib.check(self._lib.load().temp_plan_write_blob(self._handle, out.ctypes.data, _ALIGN), "temp_plan_write_blob")
        for name in self.__dict__.get("_dirty", ()):
            arr = self.__dict__[name]
            off, nb = lay[name]
            assert arr is not None and arr.nbytes == nb, "a replaced plan array must keep its size"
            out[off:off + nb] = np.ascontiguousarray(arr).view(np.uint8).reshape(-1)
        return out, lay, total


def plan_window_native(graph_dict: Dict[int, Snapshot], t_list: Sequence[int], seq_len: int, bidirectional: bool = False,
                       attention: bool = False, scan_tile: Optional[int] = None) -> WindowPlan:
    """``plan_window`` through the native planner of libtemp_b200.so (no ``transform``: the training-mode edge
    sub-sampling builds new snapshots and goes through the python planner)."""
    import ctypes as C
    from . import lib
    L = lib.load()
    (views, _keep, times), index = _snapshot_views(graph_dict)
    B = len(t_list)
    targets = (C.c_int32 * B)(*[index[int(t)] for t in t_list])
    handle = L.temp_plan_window(views, len(times), targets, B, int(seq_len), int(bidirectional), int(attention),
                                int(scan_tile or SCAN_TILE), AGG_HEAVY_DEGREE)
    if not handle:
        raise RuntimeError("temp_b200: temp_plan_window rejected its arguments")
    plan = NativeWindowPlan(handle, lib)
    cnt = lib.PlanCounts()
    lib.check(L.temp_plan_counts(handle, C.byref(cnt)), "temp_plan_counts")
    plan.seq_len, plan.batch, plan.bidirectional = int(seq_len), B, bool(bidirectional)
    plan.scan_tile = int(cnt.scan_tile)
    plan.R, plan.E = int(cnt.rows), int(cnt.edges)
    plan._n_parts = int(cnt.n_parts)
    if not bidirectional:
        plan._absent.update(("prev_b", "dt_b"))
    plan._absent.add("scan_parts" if attention else "slot_row")
    which = {name: i for i, name in enumerate(lib.PLAN_ARRAYS)}
    insts_raw = plan._fetch(which["instances"], np.int32).reshape(-1, 7).tolist()
    insts = [Instance(r[0], r[1], _DIRS[r[2]], r[3], r[4], r[5], graph_dict[times[r[6]]]) for r in insts_raw]
    for r in plan._fetch(which["segments"], np.int32).reshape(-1, 6).tolist():
        plan.segments.append(Segment(_KINDS[r[0]], r[1], r[2], r[3], insts[r[4]:r[5]]))
    for inst in plan.segments[-1].instances:
        plan.final_times.append(inst.time)
        plan.final_sizes.append(inst.n)
        plan.final_snapshots.append(inst.snapshot)
    pick = lambda a: [insts[i] if i >= 0 else None for i in a.tolist()]
    plan.last_hist_f = pick(plan._fetch(which["last_f"], np.int32))
    plan.last_hist_b = pick(plan._fetch(which["last_b"], np.int32)) if bidirectional else [None] * B
    if attention:
        plan.n_slots = int(cnt.n_slots)
        to_dicts = lambda a: [{j: insts[i] for j, i in enumerate(row) if i >= 0} for row in a.reshape(-1, B).tolist()] \
            if a.size else [dict() for _ in range(max(int(seq_len) - 1, 0))]
        plan.steps_f = to_dicts(plan._fetch(which["steps_f"], np.int32))
        plan.steps_b = to_dicts(plan._fetch(which["steps_b"], np.int32)) if bidirectional else []
    return plan


def plan_static(graph_dict: Dict[int, Snapshot], t_list: Sequence[int]) -> WindowPlan:
    """StaticRGCN (baselines/StaticRGCN.py:23-28): one snapshot per target, t_list order kept."""
    return plan_snapshots([graph_dict[int(t)] for t in t_list], t_list)


def plan_snapshots(snaps: Sequence[Snapshot], times: Sequence[int]) -> WindowPlan:
    """One step over a list of snapshots in the given order -- what ``dgl.batch(list)`` is to one encoder call
    (models/DynamicRGCN.py:92): rows = the snapshots' nodes back to back, edges in list order, no history maps.
    ``times[i]`` is the timestamp the rows of snapshot i take their time embedding from."""
    plan = WindowPlan()
    plan.seq_len, plan.batch = 1, len(snaps)
    pk = _Packer()
    seg = Segment("final", 0, 0, 0)
    for i, (snap, t) in enumerate(zip(snaps, times)):
        inst = Instance(i, 0, "c", int(t), pk.add(snap), snap.num_nodes, snap)
        pk.add_prev(snap.num_nodes, None, 0.0)
        seg.instances.append(inst)
        plan.final_times.append(int(t))
        plan.final_sizes.append(snap.num_nodes)
        plan.final_snapshots.append(snap)
    seg.row1 = pk.R
    plan.segments.append(seg)
    plan.last_hist_f = [None] * len(snaps)
    plan.last_hist_b = [None] * len(snaps)
    pk.finish(plan)
    want = np.repeat(np.asarray([int(t) for t in times], dtype=np.int32), [s.num_nodes for s in snaps]) if len(snaps) \
        else np.zeros(0, dtype=np.int32)
    if not np.array_equal(plan.row_time, want):
        plan.row_time = want
    return plan
