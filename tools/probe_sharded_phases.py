"""Development probe (N GPUs under torchrun): where the time of a snapshot-sharded forward goes -- phase 1 (RGCN layers of
the rank's snapshot block), signal / wait 1, scan of the rank's chain partitions, signal / wait 2 -- CUDA events per phase,
median over repetitions, max over ranks; GDELT-shaped config 5 at x1 (and x16 with an argument).

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/probe_sharded_phases.py [scale]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

import bench
from temp_b200.exchange import PeerGroup
from temp_b200.models import build_module
from temp_b200.snapshot import SnapshotStore

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
scale = int(sys.argv[1]) if len(sys.argv) > 1 else 1
peers = PeerGroup(dev)
store = SnapshotStore.synthetic("gdelt", num_times=24 if scale == 1 else 18, scale=scale, seed=20201116 + 4)
a = bench.make_args()
a.train_seq_len = a.test_seq_len = 15
torch.manual_seed(123)
model = build_module(a, store.num_ents, store.num_rels, store.train).to(dev).eval()
t_list = [store.times[-3], store.times[-2]]
single = model.encode(t_list)
res = model.encode_sharded(t_list, peers=peers)
fwd = res.forward
torch.cuda.synchronize()
reps = 40
marks = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(reps)]
for i in range(-5, reps):
    ev = marks[max(i, 0)]
    r = fwd.res
    ev[0].record()
    fwd._resident.run()
    ev[1].record()
    peers.barrier()
    ev[2].record()
    r.programs[1].run()
    ev[3].record()
    peers.barrier()
    ev[4].record()
torch.cuda.synchronize()
ph = np.array([[m[k].elapsed_time(m[k + 1]) for k in range(4)] for m in marks]) * 1e3
med = torch.tensor(np.median(ph, axis=0), device=dev)
tot = torch.tensor([float(np.median(ph.sum(axis=1)))], device=dev)
dist.all_reduce(med, op=dist.ReduceOp.MAX)
dist.all_reduce(tot, op=dist.ReduceOp.MAX)
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(reps):
    single.replay.run()
e.record()
torch.cuda.synchronize()
if rank == 0:
    print("x%d world %d: layers %.1f us | signal/wait 1 %.1f | scan %.1f | signal/wait 2 %.1f | total %.1f us; unsharded %.1f us" % (
        scale, world, *[float(x) for x in med.tolist()], float(tot.item()), 1e3 * s.elapsed_time(e) / reps))
    ops = [o for o in fwd._resident.ops]
    print("   phase-1 ops:", len(ops), " phase-2 ops:", len(fwd.res.programs[1].ops))
dist.destroy_process_group()
