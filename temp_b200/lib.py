"""ctypes binding of libtemp_b200.so (the C ABI declared in include/temp_b200.h).

This is the thin layer BASELINE.json's north star prescribes: PyTorch tensors in, PyTorch tensors
out, raw device pointers across the boundary.  There is NO fallback: if the shared library is
missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libtemp_b200.so")

ABI_VERSION = 20
MAX_SCAN_STEPS = 16
MAX_TERMS = 3
ACT_NONE, ACT_RELU = 0, 1
CELL_TORCH_GRU, CELL_TYPE1 = 0, 1
OP_LAYER, OP_GRU, OP_ATTN, OP_GATHER, OP_SCATTER, OP_H2D, OP_D2H, OP_GRU_SCAN = 1, 2, 3, 4, 5, 6, 7, 8

_i32 = C.c_int32
_p = C.c_void_p


class DenseTerm(C.Structure):
    _fields_ = [("a", _p), ("a_index", _p), ("a_dt", _p), ("decay_wb", _p), ("w", _p), ("w_packed", _p)]


class RgcnLayerArgs(C.Structure):
    _fields_ = [
        ("row0", _i32), ("row1", _i32), ("d", _i32),
        ("row_ptr", _p), ("e_src", _p), ("e_rel", _p), ("e_dst", _p), ("norm", _p), ("x", _p), ("weight", _p),
        ("n_bases", _i32), ("si", _i32), ("so", _i32), ("residual", _i32), ("n_terms", _i32),
        ("terms", DenseTerm * MAX_TERMS),
        ("h_bias", _p), ("activation", _i32),
        ("time_embed", _p), ("row_time", _p), ("row_time_scalar", _i32), ("te_out", _i32), ("te_chain", _i32),
        ("h_out", _p), ("chain_w", _p), ("chain_b", _p), ("chain_out", _p), ("chain_n", _i32), ("chain_ld", _i32),
        ("chain_w_packed", _p), ("inv_temperature", C.c_float), ("agg_scratch", _p), ("agg_rows", _p), ("agg_heavy", _p), ("n_agg_rows", _i32), ("n_agg_heavy", _i32),
        ("agg_lists", _i32),
        ("chain_peers", _p), ("chain_owner", _p), ("chain_world", _i32), ("reserved2", _i32),
    ]


class GruArgs(C.Structure):
    _fields_ = [
        ("row0", _i32), ("row1", _i32), ("d", _i32),
        ("gi", _p), ("gi_ld", _i32), ("gi_off", _i32),
        ("state", _p), ("prev_row", _p), ("dt", _p), ("decay_wb", _p), ("inv_temperature", C.c_float),
        ("whh_t", _p), ("whh_packed", _p), ("b_hh", _p), ("cell_type", _i32),
        ("time_embed", _p), ("row_time", _p), ("row_time_scalar", _i32),
        ("accumulate", _i32), ("out", _p), ("out_index_is_row", _i32), ("part_col", _i32), ("push", _i32),
    ]


class GruScanArgs(C.Structure):
    _fields_ = [("n_steps", _i32), ("n_parts", _i32), ("barrier", _p), ("parts", _p), ("part_stride", _i32),
                ("push_world", _i32), ("push_bufs", _p), ("push_offset", C.c_int64), ("push_row0", _i32), ("part_rows", _i32),
                ("push_multicast", _p), ("barrier_words", _i32), ("reserved3", _i32), ("steps", GruArgs * MAX_SCAN_STEPS)]


class AttnArgs(C.Structure):
    _fields_ = [
        ("row0", _i32), ("row1", _i32), ("d", _i32), ("heads", _i32),
        ("qkv", _p), ("kv_hist", _p), ("slot_row", _p), ("n_slots", _i32),
        ("tau", _p), ("decay_wb", _p), ("combine_max", _i32), ("out", _p),
    ]


class GatherArgs(C.Structure):
    _fields_ = [("n", _i32), ("d", _i32), ("table", _p), ("index", _p), ("out", _p)]


class ScatterArgs(C.Structure):
    _fields_ = [("n", _i32), ("d", _i32), ("src", _p), ("src_index", _p), ("dst_index", _p), ("add_row", _p),
                ("dst", _p)]


class ScoreLossArgs(C.Structure):
    _fields_ = [("n_pos", _i32), ("n_cand", _i32), ("d", _i32), ("score_fn", _i32), ("corrupt_tail", _i32),
                ("ent_embed", _p), ("rel_embeds", _p), ("table", _p), ("triples", _p), ("cand", _p), ("loss", _p)]


class ScoreLossBwdArgs(C.Structure):
    _fields_ = [("n_pos", _i32), ("n_cand", _i32), ("d", _i32), ("score_fn", _i32), ("corrupt_tail", _i32),
                ("ent_embed", _p), ("rel_embeds", _p), ("table", _p), ("triples", _p), ("cand", _p), ("grad_loss", _p),
                ("grad_ent_embed", _p), ("grad_rel_embeds", _p), ("grad_table", _p)]


class RankArgs(C.Structure):
    _fields_ = [("n_query", _i32), ("num_ents", _i32), ("d", _i32), ("score_fn", _i32), ("corrupt_tail", _i32),
                ("ent_embed", _p), ("rel_embeds", _p), ("table", _p), ("triples", _p), ("target", _p),
                ("filter_ptr", _p), ("filter_ids", _p), ("rank", _p)]


SCORE_FN = {"distmult": 0, "complex": 1, "transE": 2}


class SnapshotView(C.Structure):
    _fields_ = [("time", _i32), ("n_nodes", _i32), ("n_edges", _i32), ("node_ids", _p), ("row_ptr", _p), ("csr_src", _p),
                ("csr_rel", _p), ("norm", _p)]


class PlanCounts(C.Structure):
    _fields_ = [(n, _i32) for n in ("rows", "edges", "n_segments", "n_instances", "n_parts", "n_agg_rows", "n_agg_heavy",
                                    "n_slots", "batch", "seq_len", "scan_tile")]


PLAN_ARRAYS = ("ent_id", "row_time", "norm", "row_ptr", "e_src", "e_src_ent", "e_rel", "prev_a", "dt_a", "prev_b", "dt_b",
               "slot_row", "scan_parts", "agg_rows", "agg_heavy", "instances", "segments", "last_f", "last_b", "steps_f",
               "steps_b")


class CopyArgs(C.Structure):
    _fields_ = [("dst", _p), ("src", _p), ("bytes", C.c_uint64)]


class _OpUnion(C.Union):
    _fields_ = [("layer", RgcnLayerArgs), ("gru", GruArgs), ("scan", GruScanArgs), ("attn", AttnArgs), ("gather", GatherArgs),
                ("scatter", ScatterArgs), ("copy", CopyArgs)]


class Op(C.Structure):
    _fields_ = [("kind", _i32), ("reserved", _i32), ("u", _OpUnion)]


EXPORTS = ("temp_abi_version", "temp_last_error_string", "temp_device_info", "temp_rgcn_layer_fwd", "temp_rgcn_gather_fwd", "temp_gru_fwd",
           "temp_gru_scan_fwd", "temp_attention_fwd", "temp_gather_rows", "temp_scatter_rows", "temp_transpose", "temp_run_program",
           "temp_packed_weights_bytes", "temp_pack_weights", "temp_packed_gru_bytes", "temp_pack_gru_weights",
           "temp_program_kernel_count", "temp_score_loss_fwd", "temp_score_loss_bwd", "temp_rank_filtered_fwd", "temp_negative_sample", "temp_plan_window", "temp_plan_destroy", "temp_plan_counts",
           "temp_plan_array", "temp_graph_create", "temp_graph_launch", "temp_graph_destroy", "temp_peer_barrier", "temp_plan_blob_layout",
           "temp_plan_write_blob")

_lib = None


def load(path: Optional[str] = None):
    """Load the shared library (once).  Raises RuntimeError when it is missing or has the wrong ABI."""
    global _lib
    if _lib is not None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise RuntimeError("temp_b200: %s not found -- build it with `python -m temp_b200.build` "
                           "(or __graft_entry__.build()); there is no CPU fallback" % path)
    lib = C.CDLL(path)
    for name in EXPORTS:
        if not hasattr(lib, name):
            raise RuntimeError("temp_b200: %s does not export %s" % (path, name))
    lib.temp_abi_version.restype = C.c_int
    lib.temp_last_error_string.restype = C.c_char_p
    lib.temp_device_info.argtypes = [C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32)]
    lib.temp_rgcn_layer_fwd.argtypes = [C.POINTER(RgcnLayerArgs), _p]
    lib.temp_rgcn_gather_fwd.argtypes = [C.POINTER(RgcnLayerArgs), _p]
    lib.temp_gru_fwd.argtypes = [C.POINTER(GruArgs), _p]
    lib.temp_gru_scan_fwd.argtypes = [C.POINTER(GruScanArgs), _p]
    lib.temp_attention_fwd.argtypes = [C.POINTER(AttnArgs), _p]
    lib.temp_gather_rows.argtypes = [C.POINTER(GatherArgs), _p]
    lib.temp_scatter_rows.argtypes = [C.POINTER(ScatterArgs), _p]
    lib.temp_transpose.argtypes = [_p, _i32, _i32, _p, _i32, _p]
    lib.temp_run_program.argtypes = [C.POINTER(Op), _i32, _p]
    lib.temp_program_kernel_count.argtypes = [C.POINTER(Op), _i32]
    lib.temp_score_loss_fwd.argtypes = [C.POINTER(ScoreLossArgs), _p]
    lib.temp_rank_filtered_fwd.argtypes = [C.POINTER(RankArgs), _p]
    lib.temp_score_loss_bwd.argtypes = [C.POINTER(ScoreLossBwdArgs), _p]
    lib.temp_negative_sample.argtypes = [_p, _p, C.c_int64, _i32, _p, _p, _i32, _i32, _p, _p, C.c_int64]
    lib.temp_negative_sample.restype = C.c_int64
    lib.temp_peer_barrier.argtypes = [_p, _p, _i32, _i32, C.c_uint32, _p]
    lib.temp_graph_create.argtypes = [C.POINTER(Op), _i32, C.POINTER(_p)]
    lib.temp_graph_launch.argtypes = [_p, _p]
    lib.temp_graph_destroy.argtypes = [_p]
    lib.temp_plan_window.argtypes = [C.POINTER(SnapshotView), _i32, C.POINTER(_i32), _i32, _i32, _i32, _i32, _i32, _i32]
    lib.temp_plan_window.restype = _p
    lib.temp_plan_destroy.argtypes = [_p]
    lib.temp_plan_destroy.restype = None
    lib.temp_plan_counts.argtypes = [_p, C.POINTER(PlanCounts)]
    lib.temp_plan_array.argtypes = [_p, _i32, C.POINTER(C.c_int64)]
    lib.temp_plan_array.restype = _p
    lib.temp_plan_blob_layout.argtypes = [_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), _i32]
    lib.temp_plan_blob_layout.restype = C.c_int64
    lib.temp_plan_write_blob.argtypes = [_p, _p, _i32]
    lib.temp_packed_weights_bytes.argtypes = [_i32, _i32]
    lib.temp_packed_weights_bytes.restype = C.c_int64
    lib.temp_pack_weights.argtypes = [_p, _i32, _i32, _p, _p]
    lib.temp_packed_gru_bytes.argtypes = [_i32]
    lib.temp_packed_gru_bytes.restype = C.c_int64
    lib.temp_pack_gru_weights.argtypes = [_p, _i32, _p, _p]
    if lib.temp_abi_version() != ABI_VERSION:
        raise RuntimeError("temp_b200: ABI version mismatch (library %d, binding %d) -- rebuild"
                           % (lib.temp_abi_version(), ABI_VERSION))
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().temp_last_error_string().decode("utf-8", "replace")
        raise RuntimeError("temp_b200 %s failed (code %d): %s" % (what, rc, msg))


def device_info():
    sm, smem, cc = _i32(), _i32(), _i32()
    check(load().temp_device_info(C.byref(sm), C.byref(smem), C.byref(cc)), "temp_device_info")
    return {"sm_count": sm.value, "max_smem": smem.value, "cc": cc.value}


def current_stream() -> int:
    """Raw handle of torch's current stream on the current device (the private accessors skip building a Stream object:
    6 us -> 0.5 us on the cached encode path)."""
    import torch
    try:
        return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())
    except AttributeError:
        return torch.cuda.current_stream().cuda_stream


def ptr(t, byte_offset: int = 0) -> Optional[int]:
    """Device (or pinned host) address of a tensor, None for None."""
    if t is None:
        return None
    return t.data_ptr() + byte_offset


class Program(object):
    """A static launch list (TempOp[]) executed by one ``temp_run_program`` call."""

    def __init__(self):
        self.ops = []
        self._arr = None
        self.keepalive = []

    def add(self, kind: int, args) -> None:
        op = Op()
        op.kind = kind
        field = {OP_LAYER: "layer", OP_GRU: "gru", OP_GRU_SCAN: "scan", OP_ATTN: "attn", OP_GATHER: "gather", OP_SCATTER: "scatter",
                 OP_H2D: "copy", OP_D2H: "copy"}[kind]
        setattr(op.u, field, args)
        self.ops.append(op)
        self._arr = None

    def extend(self, other: "Program") -> None:
        self.ops.extend(other.ops)
        self.keepalive.extend(other.keepalive)
        self._arr = None

    def __len__(self):
        return len(self.ops)

    def fuse_gru_scans(self, barrier_ptr: int, parts_ptr: Optional[int] = None, n_parts: int = 0, part_stride: int = 0,
                       part_rows: int = 0, split_cells: bool = False, barrier_words: int = 2) -> None:
        """Replaces every run of >= 2 consecutive GRU ops by ONE scan launch (chain-partitioned when the plan's
        partition table is given, see TempGruScanArgs).  ``split_cells``: a run also ends where the recurrent cell changes
        (the Bi models: one single-cell scan per direction)."""
        out, run = [], []

        def flush():
            while run:
                chunk, rest = run[:MAX_SCAN_STEPS], run[MAX_SCAN_STEPS:]
                if len(chunk) == 1 and not split_cells:
                    out.append(chunk[0])
                else:
                    a = GruScanArgs()
                    a.n_steps, a.barrier, a.barrier_words = len(chunk), barrier_ptr, int(barrier_words)
                    if parts_ptr is not None:
                        a.parts, a.n_parts, a.part_stride, a.part_rows = parts_ptr, n_parts, part_stride, part_rows
                    for i, o in enumerate(chunk):
                        a.steps[i] = o.u.gru
                    op = Op()
                    op.kind = OP_GRU_SCAN
                    op.u.scan = a
                    out.append(op)
                run[:] = rest

        for o in self.ops:
            if o.kind == OP_GRU:
                if split_cells and run and run[-1].u.gru.b_hh != o.u.gru.b_hh:
                    flush()
                run.append(o)
            else:
                flush()
                out.append(o)
        flush()
        self.ops = out
        self._arr = None

    def count(self, kinds=(OP_LAYER, OP_GRU, OP_GRU_SCAN, OP_ATTN, OP_GATHER, OP_SCATTER)) -> int:
        return sum(1 for o in self.ops if o.kind in kinds)

    def enable_peer_push(self, bufs_dev_ptr: int, world: int, offset_elems: int, row0: int, row1: int,
                         multicast_ptr: int = 0) -> None:
        """Fused all-gather: the scan steps that write the FINAL value of rows [row0, row1) (the last op over exactly that
        range) also store it into every peer's buffer (TempGruScanArgs.push_*)."""
        # the LAST op that writes the rows must be a scan step over exactly that range (e.g. the Bi centre step of the
        # backward cell, which accumulates into the forward cell's result): walk the program backwards
        sc, last = None, None
        for o in reversed(self.ops):
            if o.kind == OP_GRU and o.u.gru.row0 < row1 and o.u.gru.row1 > row0:
                raise RuntimeError("temp_b200: the pushed rows are last written by an unfused GRU step")
            if o.kind == OP_GRU_SCAN:
                hit = [i for i in range(o.u.scan.n_steps) if o.u.scan.steps[i].row0 < row1 and o.u.scan.steps[i].row1 > row0]
                if hit:
                    st = o.u.scan.steps[hit[-1]]
                    if st.row0 != row0 or st.row1 != row1:
                        raise RuntimeError("temp_b200: no scan step covers exactly the pushed row range")
                    sc, last = o.u.scan, hit[-1]
                    break
        if sc is None:
            raise RuntimeError("temp_b200: peer push needs the fused GRU scan")
        for o in self.ops:
            if o.kind == OP_GRU_SCAN:
                for i in range(o.u.scan.n_steps):
                    o.u.scan.steps[i].push = 0
        sc.steps[last].push = 1
        sc.push_bufs, sc.push_world, sc.push_offset, sc.push_row0 = bufs_dev_ptr or None, int(world), int(offset_elems), int(row0)
        sc.push_multicast = multicast_ptr or None
        self._arr = None

    def kernel_count(self) -> int:
        """Kernel launches one ``run`` issues (a tensor-core layer with a graph part is two: gather + tile kernel)."""
        if not self.ops:
            return 0
        if self._arr is None:
            self._arr = (Op * len(self.ops))(*self.ops)
        n = load().temp_program_kernel_count(self._arr, len(self.ops))
        if n < 0:
            raise RuntimeError("temp_b200: bad program")
        return int(n)

    def capture(self) -> bool:
        """Instantiates the program as one CUDA graph (``run`` then costs a single driver call).  The program must have
        run once directly; its ops must not change afterwards.  Returns False (and keeps direct launches) when the
        driver refuses the capture."""
        if not self.ops:
            return False
        self.release_graph()
        if self._arr is None:
            self._arr = (Op * len(self.ops))(*self.ops)
        out = _p()
        rc = load().temp_graph_create(self._arr, len(self.ops), C.byref(out))
        if rc != 0 or not out.value:
            return False
        self._graph = out
        return True

    def release_graph(self) -> None:
        g = getattr(self, "_graph", None)
        if g is not None:
            load().temp_graph_destroy(g)
            self._graph = None

    def __del__(self):
        try:
            self.release_graph()
        except Exception:
            pass

    def run(self, stream: Optional[int] = None) -> None:
        if not self.ops:
            return
        st = current_stream() if stream is None else stream
        g = getattr(self, "_graph", None)
        if g is not None and self._arr is not None:
            check(load().temp_graph_launch(g, _p(st)), "temp_graph_launch")
            return
        if self._arr is None:
            self._arr = (Op * len(self.ops))(*self.ops)
        check(load().temp_run_program(self._arr, len(self.ops), _p(st)), "temp_run_program")
