"""Development probe (GPU): host time of the CACHED end-to-end call model.encode(t_list, to_host=True, reupload=True)
(the e2e arm of bench.py): wall time per call with a synchronise, host-side enqueue time alone, cProfile of the call."""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench
from temp_b200.snapshot import SnapshotStore

dev = torch.device("cuda", 0)
store = SnapshotStore.synthetic("icews14", num_times=40, scale=1, seed=bench.SEED)
model = bench.init_state(store).to(dev).eval()
t_lists = bench.batches(store, 8)
for _ in range(3):
    for tl in t_lists:
        model.encode(tl, to_host=True, reupload=True)
torch.cuda.synchronize()
n = 400
t0 = time.perf_counter()
for i in range(n):
    model.encode(t_lists[i % 8], to_host=True, reupload=True)
    torch.cuda.synchronize()
print("e2e call + synchronise: %.1f us" % (1e6 * (time.perf_counter() - t0) / n))
t0 = time.perf_counter()
for i in range(n):
    model.encode(t_lists[i % 8], to_host=True, reupload=True)
t1 = time.perf_counter()
torch.cuda.synchronize()
print("host enqueue alone: %.1f us per call (GPU-bound total %.1f us)" % (1e6 * (t1 - t0) / n, 1e6 * (time.perf_counter() - t0) / n))
t0 = time.perf_counter()
for i in range(n):
    torch.cuda.synchronize()
print("empty synchronise: %.1f us" % (1e6 * (time.perf_counter() - t0) / n))
pr = cProfile.Profile()
pr.enable()
for i in range(n):
    model.encode(t_lists[i % 8], to_host=True, reupload=True)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(25)
