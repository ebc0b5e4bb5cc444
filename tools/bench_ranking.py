"""Development probe (GPU): filtered ranking of one target graph, CUDA path vs the torch formulation.

    python tools/bench_ranking.py [n_query]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import json

import numpy as np
import torch

import bench
from temp_b200.evaluation import EvaluationFilter
from temp_b200.snapshot import SnapshotStore

dev = torch.device("cuda", 0)
store = SnapshotStore.synthetic("icews14", num_times=40, scale=1, seed=bench.SEED)
model = bench.init_state(store).to(dev).eval()
tl = bench.batches(store, 1)[0]
res = model.encode(tl)
i = 0
t = res.plan.final_times[i]
g = store.train[t]
table = model.all_embeds(res, i)
src, dst = g.edges()
samples = torch.stack([src, g.edata["type_s"], dst]).transpose(0, 1).to(dev)
if len(sys.argv) > 1:
    samples = samples[:int(sys.argv[1])]
ev = EvaluationFilter(model.args, model.calc_score, store.train, {}, {})
ent = res.per_graph[i]


def timed(f, n=20):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        f()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


a = ev.calc_metrics_single_graph(ent, model.rel_embeds, table, samples, g, t)
b = ev.calc_metrics_single_graph_torch(ent, model.rel_embeds, table, samples, g, t)
fast = timed(lambda: ev.calc_metrics_single_graph(ent, model.rel_embeds, table, samples, g, t))
slow = timed(lambda: ev.calc_metrics_single_graph_torch(ent, model.rel_embeds, table, samples, g, t))
# kernels alone: the filter lists are host work of both routes
import ctypes as C
from temp_b200 import lib
ids = torch.from_numpy(g.node_ids).to(dev)
ptr, flat = ev.filter_lists(samples.cpu(), t, g, "tail")
ptr_d, flat_d = torch.from_numpy(ptr).to(dev), torch.from_numpy(flat).to(dev)
target = ids[samples[:, 2]].contiguous()
out = torch.empty(samples.shape[0], dtype=torch.long, device=dev)
args = lib.RankArgs(samples.shape[0], table.shape[0], table.shape[1], lib.SCORE_FN["complex"], 1, ent.data_ptr(),
                    model.rel_embeds.data_ptr(), table.data_ptr(), samples.data_ptr(), target.data_ptr(), ptr_d.data_ptr(),
                    flat_d.data_ptr(), out.data_ptr())
L = lib.load()
kern = timed(lambda: L.temp_rank_filtered_fwd(C.byref(args), C.c_void_p(lib.current_stream())), 200)
print(json.dumps({"queries": int(samples.shape[0]), "entities": int(table.shape[0]), "d": int(table.shape[1]),
                  "equal_ranks": int((a == b).sum()), "of": int(a.numel()),
                  "cuda_route_ms": round(fast, 4), "torch_route_ms": round(slow, 4),
                  "rank_kernel_one_side_us": round(kern * 1e3, 2),
                  "table_bytes": int(table.numel() * 4)}))
