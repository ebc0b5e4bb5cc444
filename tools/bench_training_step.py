"""Development probe (GPU): one main.py-style training step (train() mode forward + backward + Adam) on the bench
workload's shapes, with the fused scorer (temp_score_loss_fwd / _bwd) and with the reference's materialised torch
formulation.  The encoder runs through the torch autograd fallback in both (its kernels are forward-only).

    python tools/bench_training_step.py [negative_rate]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from temp_b200.snapshot import SnapshotStore

neg = int(sys.argv[1]) if len(sys.argv) > 1 else 500
dev = torch.device("cuda", 0)
store = SnapshotStore.synthetic("icews14", num_times=40, scale=1, seed=bench.SEED)
out = {"negative_rate": neg}
for fused in (True, False):
    model = bench.init_state(store).to(dev)
    model.args.negative_rate = model.negative_rate = neg
    model._corrupter = None
    model.fused_scorer = fused
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    tl = bench.batches(store, 1)[0]

    def step():
        opt.zero_grad()
        loss = model.forward(torch.tensor(tl))
        loss.backward()
        opt.step()
        return float(loss.detach())

    np.random.seed(0)
    torch.manual_seed(0)
    for _ in range(2):
        first = step()
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats()
    t0 = time.perf_counter()
    n = 5
    for _ in range(n):
        last = step()
    torch.cuda.synchronize()
    out["fused_scorer" if fused else "torch_scorer"] = {"s_per_step": (time.perf_counter() - t0) / n, "loss": last,
                                                        "peak_bytes": int(torch.cuda.max_memory_allocated())}
print(json.dumps(out))
if os.environ.get("TEMP_PROFILE"):
    import cProfile
    import pstats
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
