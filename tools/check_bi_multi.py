"""torchrun script (one process per GPU): the BIDIRECTIONAL GRU model through both multi-GPU paths -- the snapshot-sharded
forward with in-kernel exchanges against the unsharded one (bit-identical), and the fused all-gather of the final states
(model.encode(exchange=...)) against NCCL -- on window batches that include the last timestamps (no backward history: the
backward chain is the centre step alone).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_bi_multi.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

import bench
from temp_b200.exchange import FinalStateAllGather, PeerGroup
from temp_b200.models import build_module
from temp_b200.snapshot import SnapshotStore

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)

store = SnapshotStore.synthetic("icews14", num_times=24, scale=1, seed=bench.SEED)
args = bench.make_args("BiGRRGCN")
torch.manual_seed(123)
model = build_module(args, store.num_ents, store.num_rels, store.train).to(dev).eval()
peers = PeerGroup(dev)
times = store.times
batches = [[times[10], times[11], times[12]], [times[-1]], [times[-2], times[-1]], [times[0], times[5]]]


def all_equal(flag, what):
    f = torch.tensor([int(flag)], device=dev)
    dist.all_reduce(f, op=dist.ReduceOp.MIN)
    if int(f.item()) != 1:
        raise SystemExit("rank %d: %s" % (rank, what))


for tl in batches:
    want = model.encode(tl).out.clone()
    res = model.encode_sharded(tl, peers=peers)
    torch.cuda.synchronize()
    all_equal(torch.equal(res.out, want), "sharded Bi forward differs for %r (max %.3e)" % (tl, float((res.out - want).abs().max())))

# fused all-gather: every rank encodes its own batch (rank r takes batch r), all ranks end up with all final states
mine = batches[rank % len(batches)]
rows = torch.tensor([max(model.plan(b).final.row1 - model.plan(b).final.row0 for b in batches)], device=dev)
ex = FinalStateAllGather(dev, int(rows.item()), model.embed_size)
for _ in range(3):
    res = model.encode(mine, exchange=ex)
torch.cuda.synchronize()
nf = res.out.shape[0]
sizes = [torch.zeros(1, dtype=torch.long, device=dev) for _ in range(world)]
dist.all_gather(sizes, torch.tensor([nf], device=dev))
pad = torch.zeros(int(rows.item()), model.embed_size, device=dev)
pad[:nf] = res.out
got = [torch.zeros_like(pad) for _ in range(world)]
dist.all_gather(got, pad)
ok = all(torch.equal(ex.gathered(k, int(sizes[k].item())), got[k][:int(sizes[k].item())]) for k in range(world))
all_equal(ok, "fused all-gather of the Bi final states differs from NCCL")
if rank == 0:
    print("bi multi ok: %d batches sharded bit-identically; fused all-gather (%s) equals NCCL" % (len(batches), "fused" if ex.fused else "nccl"))
dist.destroy_process_group()
