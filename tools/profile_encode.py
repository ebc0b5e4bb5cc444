"""Development probe (GPU): where the host time of an uncached model.encode(t_list) goes (cProfile, bench workload)."""
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
model.encode_cache_size = 0
for tl in t_lists:
    model.encode(tl)
torch.cuda.synchronize()
n = 200
t0 = time.perf_counter()
for i in range(n):
    model.encode(t_lists[i % 8])
torch.cuda.synchronize()
print("uncached encode: %.3f ms per call" % (1e3 * (time.perf_counter() - t0) / n))
t0 = time.perf_counter()
for i in range(n):
    model.plan(t_lists[i % 8])
print("plan only: %.3f ms per call" % (1e3 * (time.perf_counter() - t0) / n))
plans = [model.plan(tl) for tl in t_lists]
t0 = time.perf_counter()
for i in range(n):
    model.runtime.build(plans[i % 8])
print("build only: %.3f ms per call" % (1e3 * (time.perf_counter() - t0) / n))
pr = cProfile.Profile()
pr.enable()
for i in range(n):
    model.encode(t_lists[i % 8], to_host=True)
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)
