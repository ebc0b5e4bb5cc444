"""ncu target (GPU): replays the forward of one BASELINE configuration (tests/test_gpu_fullsize.CONFIGS, full size, synthetic
snapshots) a few times, nothing else.

    ncu --set full --clock-control none --import-source on -k regex:'tcw|gather_wide' -s 9 -c 9 \
        -o gpurun_out/r2_full_config3 python tools/ncu_config.py config3_bigrrgcn_icews0515_nb100 3
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tests import test_gpu_fullsize as T

name = sys.argv[1] if len(sys.argv) > 1 else "config3_bigrrgcn_icews0515_nb100"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cfg = next(c for c in T.CONFIGS if c[0] == name)
model, _, t_list = T._build(cfg)
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
res = model.encode(t_list)
torch.cuda.synchronize()
for _ in range(reps):
    flush.fill_(1.0)
    res.program.run()
torch.cuda.synchronize()
print("rows %d edges %d launches %d" % (res.plan.R, res.plan.E, res.program.kernel_count()))
