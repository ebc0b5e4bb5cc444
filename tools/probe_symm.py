"""Probe (2+ GPUs): torch symmetric memory gives peer-mapped device pointers usable from our own kernels?"""
import os, sys, time
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
t = symm.empty((world, 1024, 128), dtype=torch.float32, device=dev)
t.zero_()
hdl = symm.rendezvous(t, group=dist.group.WORLD)
print(rank, "rendezvous ok", type(hdl).__name__, [a for a in dir(hdl) if not a.startswith("_")][:40])
bufs = [hdl.get_buffer(r, (world, 1024, 128), torch.float32) for r in range(world)]
print(rank, "peer ptrs", [hex(b.data_ptr()) for b in bufs])
hdl.barrier()
# push my slab into every peer with plain torch copies on peer-mapped tensors
mine = torch.full((1024, 128), float(rank + 1), device=dev)
for r in range(world):
    bufs[r][rank].copy_(mine)
hdl.barrier()
torch.cuda.synchronize()
ok = all(float(t[r].mean()) == r + 1 for r in range(world))
print(rank, "all-gather by peer stores ok:", ok)
# timing: barrier cost
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(50):
    hdl.barrier()
e.record(); torch.cuda.synchronize()
print(rank, "symm barrier us", s.elapsed_time(e) / 50 * 1e3)
send = torch.zeros(1024, 128, device=dev); recv = torch.empty(world * 1024, 128, device=dev)
for _ in range(5): dist.all_gather_into_tensor(recv, send)
torch.cuda.synchronize(); s.record()
for _ in range(50): dist.all_gather_into_tensor(recv, send)
e.record(); torch.cuda.synchronize()
print(rank, "nccl all_gather (512 KB/rank) us", s.elapsed_time(e) / 50 * 1e3)
dist.destroy_process_group()
