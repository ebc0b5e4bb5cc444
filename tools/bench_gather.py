"""Development probe (GPU): the aggregation launch alone (temp_rgcn_gather_fwd), LDG variant against the TMA bulk-copy
variant (TEMP_GATHER=ldg|bulk, one process each), on the layer-2 launch of one window batch.

    TEMP_GATHER=bulk python tools/bench_gather.py icews14 16
Prints one JSON line: time per launch (CUDA events, L2 flushed), compulsory bytes (unique source rows once + indices +
aggregate rows written), edge-streamed bytes E * (4 D + 8), a checksum of the aggregates (variants must agree bit for bit)."""
import ctypes as C
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import bench
from temp_b200 import lib
from temp_b200.snapshot import SnapshotStore

shape = sys.argv[1] if len(sys.argv) > 1 else "icews14"
scale = int(sys.argv[2]) if len(sys.argv) > 2 else 16
dev = torch.device("cuda", 0)
bench.WORKLOAD.update(shape=shape, seq_len=15 if shape == "gdelt" else 8, batch=2 if shape == "gdelt" else 8)
store = SnapshotStore.synthetic(shape, num_times=18, scale=scale, seed=bench.SEED)
model = bench.init_state(store).to(dev).eval()
tl = bench.batches(store, 1)[0]
res = model.encode(tl)
torch.cuda.synchronize()
layer_ops = [o for o in res.program.ops if o.kind == lib.OP_LAYER]
a = layer_ops[-1].u.layer                      # layer 2: x = h1 (device rows), agg_scratch
plan = res.plan
D = 128
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
L = lib.load()
st = C.c_void_p(lib.current_stream())
agg = res.runtime_agg = model.runtime.ws.get("agg", plan.R * D)
agg.zero_()
lib.check(L.temp_rgcn_gather_fwd(C.byref(a), st), "gather")
torch.cuda.synchronize()
digest = hashlib.sha1(agg[:plan.R * D].cpu().numpy().tobytes()).hexdigest()[:16]
# fp64 statement of the aggregation (models/RGCN.py:91-104, 1x1 blocks) on the same inputs
h1 = res.bufs["h1"].double()
rp = plan.row_ptr.astype(np.int64)
dst = torch.as_tensor(np.repeat(np.arange(plan.R), np.diff(rp)), device=dev).long()
src_t = torch.as_tensor(plan.e_src, device=dev).long()
rel_t = torch.as_tensor(plan.e_rel, device=dev).long()
nrm_t = torch.as_tensor(plan.norm, device=dev).double()
W2 = model.ent_encoder.layer_2.weight.detach().double()
want = torch.zeros(plan.R, D, dtype=torch.float64, device=dev).index_add_(0, dst, h1[src_t] * W2[rel_t] * nrm_t[dst][:, None]) * nrm_t[:, None]
err = float((agg[:plan.R * D].view(plan.R, D).double() - want).abs().max() / want.abs().max())
reps = 20
s = [torch.cuda.Event(enable_timing=True) for _ in range(reps)]
e = [torch.cuda.Event(enable_timing=True) for _ in range(reps)]
for i in range(reps):
    flush.fill_(1.0)
    s[i].record()
    L.temp_rgcn_gather_fwd(C.byref(a), st)
    e[i].record()
torch.cuda.synchronize()
ms = float(np.median([x.elapsed_time(y) for x, y in zip(s, e)]))
deg = np.diff(plan.row_ptr)
nz = int((deg > 0).sum())
uniq_src = int(np.unique(plan.e_src).shape[0])
rel_rows = int(np.unique(plan.e_rel).shape[0])
compulsory = uniq_src * 4 * D + plan.E * 8 + nz * (4 * D + 16) + rel_rows * 4 * D
streamed = plan.E * (4 * D + 8) + nz * (4 * D + 16)
peak = 6543.1
print(json.dumps({"variant": os.environ.get("TEMP_GATHER", "ldg"), "shape": shape, "scale": scale, "rows": int(plan.R), "edges": int(plan.E),
                  "rows_with_in_edges": nz, "max_in_degree": int(deg.max()), "ms": ms, "compulsory_bytes": compulsory,
                  "edge_streamed_bytes": streamed, "compulsory_GBps": compulsory / ms / 1e6, "streamed_GBps": streamed / ms / 1e6,
                  "frac_compulsory_of_measured_hbm": compulsory / ms / 1e6 / peak, "frac_streamed": streamed / ms / 1e6 / peak,
                  "agg_sha1": digest, "max_err_over_scale_vs_fp64": err}))
