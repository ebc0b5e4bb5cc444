"""ncu target (GPU): replays the bench forward of one window batch at a given scale, nothing else.

    ncu --set full --clock-control none --import-source on -k regex:'rgcn_gather|rgcn_layer_tc|gru_scan_tm' -s 12 -c 4 \
        -o gpurun_out/r2_full_x16 python tools/ncu_forward.py 16 6
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from temp_b200.snapshot import SnapshotStore

scale = int(sys.argv[1]) if len(sys.argv) > 1 else 1
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
dev = torch.device("cuda", 0)
store = SnapshotStore.synthetic("icews14", num_times=40 if scale == 1 else 16, scale=scale, seed=bench.SEED)
model = bench.init_state(store).to(dev).eval()
tl = bench.batches(store, 1)[0]
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
res = model.encode(tl)
for _ in range(reps):
    flush.fill_(1.0)
    res.replay.run()
torch.cuda.synchronize()
print("rows %d edges %d" % (res.plan.R, res.plan.E))
