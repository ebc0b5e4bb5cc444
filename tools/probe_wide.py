"""Development probe for the 64-row tcgen05 layer (tc_wide.cu) on BASELINE config 3 as shipped (BiGRRGCN, D = 200,
n_bases = 100): whole-forward device time and the time of every launch-program op alone (CUDA events, L2 flushed, median
of 20), error against the oracle.  Run once per path:  TEMP_WIDE_TC=0 python tools/probe_wide.py  (fp32 SIMT layers)
and  python tools/probe_wide.py  (tensor-core layers).  One JSON line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from temp_b200 import lib
from tests import test_gpu_fullsize as T

name = sys.argv[1] if len(sys.argv) > 1 else "config3_bigrrgcn_icews0515_nb100"
cfg = next(c for c in T.CONFIGS if c[0] == name)
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
model, oracle, t_list = T._build(cfg)
res = model.encode(t_list)
torch.cuda.synchronize()
with torch.no_grad():
    want = torch.cat(oracle.evaluate_embed(t_list)["per_graph"]).numpy()
got = res.out.cpu().numpy()
err = float(np.abs(got.astype(np.float64) - want).max() / np.abs(want).max())


def timed(prog, n=20):
    ts = []
    for _ in range(n):
        flush.fill_(1.0)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        prog.run()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return float(np.median(ts))


total = timed(res.program)
kinds = {lib.OP_LAYER: "layer", lib.OP_GRU: "gru", lib.OP_GRU_SCAN: "scan", lib.OP_ATTN: "attn", lib.OP_H2D: "h2d"}
ops = []
for o in res.program.ops:
    pr = lib.Program()
    pr.ops = [o]
    d = {"kind": kinds.get(o.kind, str(o.kind)), "ms": timed(pr, 10)}
    if o.kind == lib.OP_LAYER:
        d.update(rows=o.u.layer.row1 - o.u.layer.row0, chain_n=o.u.layer.chain_n, launches=pr.kernel_count())
    if o.kind == lib.OP_GRU_SCAN:
        d.update(steps=o.u.scan.n_steps)
    ops.append(d)
print(json.dumps({"config": name, "wide_tc": os.environ.get("TEMP_WIDE_TC", "1") != "0", "rows": int(res.plan.R), "edges": int(res.plan.E),
                  "forward_ms": total, "kernel_launches": res.program.kernel_count(), "err_vs_oracle": err, "ops": ops}))
