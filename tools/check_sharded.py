"""torchrun script (one process per GPU): the snapshot-sharded forward (temp_b200/sharding.py) against the unsharded one
on the same inputs, then a timing of both on a GDELT-shaped batch (BASELINE.json config 5: GRRGCN, seq_len 15, B = 2).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/check_sharded.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

import bench
from temp_b200.snapshot import SnapshotStore

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)

scale = int(sys.argv[1]) if len(sys.argv) > 1 else 1
store = SnapshotStore.synthetic("gdelt", num_times=24, scale=scale, seed=20201116 + 4)
bench.WORKLOAD.update(seq_len=15)
model = bench.init_state(store).to(dev).eval()
t_list = [store.times[20], store.times[21]]
want = model.encode(t_list).out.clone()
res = model.encode_sharded(t_list)
torch.cuda.synchronize()
same = torch.equal(res.out, want)
flags = torch.tensor([int(same)], device=dev)
dist.all_reduce(flags, op=dist.ReduceOp.MIN)
if int(flags.item()) != 1:
    raise SystemExit("rank %d: sharded forward differs from the unsharded one (max abs diff %.3e)"
                     % (rank, float((res.out - want).abs().max())))


def timed(fn, reps=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    t = torch.tensor([s.elapsed_time(e) / reps], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


single = model.encode(t_list)
ms_single = timed(lambda: single.program.run())
ms_shard = timed(lambda: model.encode_sharded(prepared=res))
if rank == 0:
    print("sharded ok")
    print(json.dumps({"workload": "GRRGCN rec-only-last-layer, GDELT-shaped synthetic x%d, seq_len=15, B=2" % scale,
                      "n_gpus": world, "rows": int(res.plan.R), "edges": int(res.plan.E),
                      "ms_unsharded_one_gpu": ms_single, "ms_snapshot_sharded": ms_shard,
                      "edges_per_s_sharded": res.plan.E / (ms_shard * 1e-3),
                      "row_blocks": [int(x) for x in np.diff(res.shard.row_bounds)],
                      "partitions_per_rank": [int(x) for x in np.diff(res.shard.part_offset)]}))
dist.destroy_process_group()
