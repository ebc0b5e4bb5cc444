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


def check(res, what):
    torch.cuda.synchronize()
    flags = torch.tensor([int(torch.equal(res.out, want))], device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    if int(flags.item()) != 1:
        raise SystemExit("rank %d: %s sharded forward differs from the unsharded one (max abs diff %.3e)"
                         % (rank, what, float((res.out - want).abs().max())))


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


# transport 1: torch.distributed collectives between the launches (NCCL)
res = model.encode_sharded(t_list)
check(res, "collective-transport")
ms_coll = timed(lambda: model.encode_sharded(prepared=res))
# transport 2: in-kernel NVLink peer stores over symmetric memory
from temp_b200.exchange import PeerGroup
peers = PeerGroup(dev)
res2 = model.encode_sharded(t_list, peers=peers)
check(res2, "peer-store")
for _ in range(3):                       # re-runs of the prepared forward stay identical
    model.encode_sharded(prepared=res2)
    check(res2, "peer-store (re-run)")
ms_peer = timed(lambda: model.encode_sharded(prepared=res2))
single = model.encode(t_list)
ms_single = timed(lambda: single.replay.run())
if rank == 0:
    print("sharded ok")
    print(json.dumps({"world": world, "scale": scale, "rows": int(res.plan.R), "edges": int(res.plan.E),
                      "ms_unsharded_one_gpu": ms_single, "ms_sharded_peer_stores": ms_peer, "ms_sharded_collectives": ms_coll}))
dist.destroy_process_group()
