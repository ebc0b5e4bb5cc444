"""Snapshot-sharded execution of one window batch over the GPUs of a box.

The reference's only parallelism is data parallel over target timestamps (``DistributedSampler``,
models/TKG_Module.py:166-168); ``bench.py --gpus N`` keeps that as the headline (weak scaling).  This module adds the
second axis the path offers -- the one BASELINE.json's config 5 names -- for the ``--rec-only-last-layer`` GRU
families: every recurrence-free launch (layer 1, layer-2 aggregation + self loop + GRU input gates;
models/RRGCN.py:198-200) is independent per snapshot instance, and the GRU scan is independent per chain partition
(temp_b200/planner.py), so ONE window batch is cut twice:

  phase 1  rank r runs both RGCN layers for a contiguous block of snapshot instances (balanced by rows + edges)
           -> exchange 1: all ranks receive every block's GRU input pre-activations ``gi`` (one broadcast per block)
  phase 2  rank r scans its share of the chain partitions (round-robin over the size-sorted table)
           -> exchange 2: NCCL all-gather of the final-layer entity states of the target snapshots
              (the collective the north star names "before scoring")

Two transports.  On the GPUs of one box (``PeerShardedForward``, symmetric memory over NVLink) both exchanges are
IN-KERNEL peer stores and no collective call is on the data path: exchange 1 happens in the tile kernel's chain epilogue
-- every ``gi`` row is written straight into the memory of the ONE rank that will scan the row's chain partition
(``TempRgcnLayerArgs.chain_peers / chain_owner``: 1/world of the round-1 traffic, which broadcast all of ``gi`` to every
rank) -- and exchange 2 in the scan kernel's final step, which stores each final state into every rank's state buffer at
its own packed row (``TempGruScanArgs.push_*``; no pack / unpack).  One signal / wait launch (``temp_peer_barrier``)
follows each phase.  The portable transport (``exchange_blocks`` / ``exchange_rows``: ``torch.distributed`` collectives,
NCCL or gloo) remains as the fallback and for the CPU tests.  Every rank plans the whole batch (planning is deterministic
and cheap) and holds full-size buffers; only the work is sharded.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np
import torch

from .planner import WindowPlan

__all__ = ["ShardPlan", "make_shard_plan", "exchange_blocks", "exchange_rows", "PeerShardedForward"]


@dataclass
class ShardPlan:
    world: int
    row_bounds: np.ndarray            # [world + 1] packed-row block of every rank (cut at instance boundaries)
    part_offset: np.ndarray           # [world + 1] range of every rank in the rank-major chain-partition table
    parts: np.ndarray                 # [P, n_seg, 2] the plan's partition table reordered rank-major
    final_rows: List[np.ndarray]      # per rank: packed rows of the final segment whose state the rank computes
    cache: dict = None                # device-side index tensors / slabs of exchange_rows, per plan
    row_owner: np.ndarray = None      # [R] int32: the rank whose chain partitions hold the packed row (its scan reads the row's gi)

    def rows_of(self, rank: int):
        return int(self.row_bounds[rank]), int(self.row_bounds[rank + 1])

    def parts_of(self, rank: int):
        return int(self.part_offset[rank]), int(self.part_offset[rank + 1])


def make_shard_plan(plan: WindowPlan, world: int) -> ShardPlan:
    """Cuts ``plan`` for ``world`` ranks (deterministic: every rank computes the same cut)."""
    # ---- phase 1: contiguous blocks of snapshot instances, balanced by a rows + edges cost ------------------------
    insts = [inst for seg in plan.segments for inst in seg.instances]      # packed order
    cost = np.array([inst.n + inst.snapshot.num_edges for inst in insts], dtype=np.float64)
    ends = np.array([inst.row0 + inst.n for inst in insts], dtype=np.int64)
    cum = np.cumsum(cost)
    bounds = [0]
    for r in range(1, world):
        target = cum[-1] * r / world
        k = int(np.searchsorted(cum, target, side="left"))                 # block r-1 ends with instance k
        k = min(max(k, 0), len(insts) - 1)
        bounds.append(max(int(ends[k]), bounds[-1]))
    bounds.append(int(plan.R))
    # ---- phase 2: chain partitions, size-sorted table dealt round-robin, stored rank-major ------------------------
    table = plan.scan_parts
    order = [np.arange(r, table.shape[0], world) for r in range(world)]
    part_offset = np.zeros(world + 1, dtype=np.int64)
    np.cumsum([len(o) for o in order], out=part_offset[1:])
    parts = np.ascontiguousarray(np.concatenate([table[o] for o in order], axis=0)) if table.shape[0] else table
    fin = len(plan.segments) - 1
    final_rows = []
    for r in range(world):
        rng = parts[part_offset[r]:part_offset[r + 1], fin]
        final_rows.append(np.concatenate([np.arange(lo, hi, dtype=np.int64) for lo, hi in rng] or
                                         [np.zeros(0, dtype=np.int64)]))
    owner = np.zeros(int(plan.R), dtype=np.int32)
    for r in range(world):
        for lo, hi in parts[part_offset[r]:part_offset[r + 1]].reshape(-1, 2):
            owner[lo:hi] = r
    return ShardPlan(world, np.asarray(bounds, dtype=np.int64), part_offset, parts, final_rows, {}, owner)


def exchange_blocks(buf: torch.Tensor, row_bounds: Sequence[int], group=None) -> None:
    """Exchange 1: ``buf[row_bounds[r]:row_bounds[r+1]]`` is valid on rank r; afterwards every rank holds all blocks.
    One in-place broadcast per non-empty block (exact sizes, no staging copies)."""
    import torch.distributed as dist
    for r in range(len(row_bounds) - 1):
        lo, hi = int(row_bounds[r]), int(row_bounds[r + 1])
        if hi > lo:
            dist.broadcast(buf[lo:hi], src=dist.get_global_rank(group, r) if group is not None else r, group=group)


def exchange_rows(state: torch.Tensor, rows_per_rank: Sequence[np.ndarray], rank: int, group=None,
                  cache: Optional[dict] = None) -> None:
    """Exchange 2: rank r holds valid ``state[rows_per_rank[r]]``; an all-gather of the padded row slabs makes every
    listed row valid on every rank (the final-layer entity states before scoring)."""
    import torch.distributed as dist
    world = len(rows_per_rank)
    width = max((len(x) for x in rows_per_rank), default=0)
    if width == 0:
        return
    key = ("exchange_rows", str(state.device), int(state.shape[1]))   # the cache belongs to ONE shard plan
    if cache is not None and key in cache:
        mine, dst_rows, src_slots, send, recv = cache[key]
    else:
        mine = torch.as_tensor(rows_per_rank[rank], dtype=torch.long, device=state.device)
        dst_rows = torch.as_tensor(np.concatenate(rows_per_rank), dtype=torch.long, device=state.device)
        src_slots = torch.as_tensor(np.concatenate([r * width + np.arange(len(x)) for r, x in enumerate(rows_per_rank)]),
                                    dtype=torch.long, device=state.device)
        send = torch.zeros(width, state.shape[1], dtype=state.dtype, device=state.device)
        recv = torch.empty(world * width, state.shape[1], dtype=state.dtype, device=state.device)
        if cache is not None:
            cache[key] = (mine, dst_rows, src_slots, send, recv)
    if mine.numel():
        send[:mine.numel()] = state.index_select(0, mine)
    dist.all_gather_into_tensor(recv, send, group=group)
    state.index_copy_(0, dst_rows, recv.index_select(0, src_slots))


class PeerShardedForward(object):
    """One window batch cut over the GPUs of a box with in-kernel NVLink exchanges (module docstring).  Collective: every
    rank constructs it with the same batch; ``run()`` on every rank leaves the complete final-layer states in
    ``res.out`` everywhere."""

    def __init__(self, model, plan: WindowPlan, peers):
        import ctypes as C
        from . import lib
        self.model, self.plan, self.peers = model, plan, peers
        rt = model.runtime
        rank, world = peers.rank, peers.world
        D = model.embed_size
        shard = make_shard_plan(plan, world)
        G = 3 * D * (2 if plan.bidirectional else 1)
        # symmetric gi / state buffers, sized by the batch (identical on every rank: all ranks plan the whole batch)
        cap_gi, cap_s = int(plan.R) * G, int(plan.R) * D
        sym = getattr(rt, "_sym", None)
        if sym is None or sym["gi"][0].numel() < cap_gi or sym["state"][0].numel() < cap_s or sym["peers"] is not peers:
            sym = rt._sym = {"peers": peers, "gi": peers.alloc((int(cap_gi * 1.25) + 64,)),
                             "state": peers.alloc((int(cap_s * 1.25) + 64,))}
        rt.ws._bufs["gi_l2"], rt.ws._bufs["state"] = sym["gi"][0], sym["state"][0]
        res = rt.build_sharded(plan, shard, rank)
        owner = torch.as_tensor(shard.row_owner, dtype=torch.int32, device=rt.device)
        for o in res.programs[0].ops:                 # exchange 1: the chained output goes to the scanning rank
            if o.kind == lib.OP_LAYER and o.u.layer.chain_out:
                if o.u.layer.chain_out != sym["gi"][0].data_ptr():
                    raise RuntimeError("temp_b200: the sharded forward expects the chained gi buffer to be symmetric")
                o.u.layer.chain_peers = sym["gi"][1].data_ptr()
                o.u.layer.chain_owner = owner.data_ptr()
                o.u.layer.chain_world = world
        res.programs[0]._arr = None
        fin = plan.final                              # exchange 2: final states into every rank's state buffer
        scans = [o for o in res.programs[1].ops if o.kind == lib.OP_GRU_SCAN]
        if not scans:
            raise RuntimeError("temp_b200: the sharded forward expects the fused scan launch")
        sc = scans[-1].u.scan                        # (the Bi models run one scan per direction: the last one writes the result)
        for i in range(sc.n_steps):
            sc.steps[i].push = 1 if (sc.steps[i].row0 == fin.row0 and sc.steps[i].row1 == fin.row1 and
                                      i == max(k for k in range(sc.n_steps)
                                               if sc.steps[k].row0 == fin.row0 and sc.steps[k].row1 == fin.row1)) else 0
        sc.push_bufs, sc.push_world, sc.push_offset, sc.push_row0 = sym["state"][1].data_ptr(), world, 0, 0
        res.programs[1]._arr = None
        res.programs[0].keepalive += [owner, sym["gi"][0], sym["state"][0]]
        res.shard, res.sharded = shard, True
        self.res = res
        self._resident = lib.Program()                # phase 1 without the plan upload (the plan stays on the device)
        self._resident.ops = [o for o in res.programs[0].ops if o.kind != lib.OP_H2D]
        self._resident.keepalive = res.programs[0].keepalive
        self._uploaded = False

    def run(self):
        res = self.res
        (self._resident if self._uploaded else res.programs[0]).run()
        self._uploaded = True
        self.peers.barrier()          # every rank's gi rows have landed where they will be scanned
        res.programs[1].run()
        self.peers.barrier()          # every rank's final states have landed everywhere
        return res
