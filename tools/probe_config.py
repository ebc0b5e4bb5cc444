"""Timing probe (GPU): per-op breakdown of the forward of one BASELINE configuration (index into
tests/test_gpu_fullsize.py CONFIGS), L2 flushed and warm.

    python tools/probe_config.py 2        # config 3: BiGRRGCN, ICEWS05-15 shape, D = 200, n_bases = 100 (fp32 SIMT path)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from temp_b200 import lib
from tests.test_gpu_fullsize import CONFIGS, _build

idx = int(sys.argv[1]) if len(sys.argv) > 1 else 2
model, _, t_list = _build(CONFIGS[idx])
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")


def time_prog(prog, reps=20, do_flush=True):
    s = [torch.cuda.Event(enable_timing=True) for _ in range(reps)]
    e = [torch.cuda.Event(enable_timing=True) for _ in range(reps)]
    for i in range(reps):
        if do_flush:
            flush.fill_(1.0)
        s[i].record()
        prog.run()
        e[i].record()
    torch.cuda.synchronize()
    return float(np.median([a.elapsed_time(b) for a, b in zip(s, e)])) * 1e3


res = model.encode(t_list)
torch.cuda.synchronize()
ops = [o for o in res.program.ops if o.kind != lib.OP_H2D]
full = lib.Program()
full.ops = ops
print("%s rows=%d edges=%d ops=%d kernels=%d  forward: flush %.1f us, warm %.1f us" % (
    CONFIGS[idx][0], res.plan.R, res.plan.E, len(ops), full.kernel_count(), time_prog(full), time_prog(full, do_flush=False)))
names = {1: "layer", 2: "gru", 3: "attn", 4: "gather", 5: "scatter", 8: "scan"}
for i, o in enumerate(ops):
    one = lib.Program()
    one.ops = [o]
    extra = ""
    if o.kind == 1:
        extra = "rows=%d chain_n=%d terms=%d" % (o.u.layer.row1 - o.u.layer.row0, o.u.layer.chain_n, o.u.layer.n_terms)
    if o.kind == 2:
        extra = "rows=%d" % (o.u.gru.row1 - o.u.gru.row0)
    if o.kind == 8:
        extra = "steps=%d" % o.u.scan.n_steps
    print("   op %d %-6s %-28s flush %.1f us  warm %.1f us" % (i, names.get(o.kind, o.kind), extra, time_prog(one), time_prog(one, do_flush=False)))
