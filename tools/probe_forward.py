"""Timing probe (GPU): per-op breakdown of the bench forward, with and without L2 flush."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from temp_b200 import lib
from temp_b200.snapshot import SnapshotStore

dev = torch.device("cuda", 0)
scale = int(sys.argv[1]) if len(sys.argv) > 1 else 1
store = SnapshotStore.synthetic("icews14", num_times=40 if scale == 1 else 16, scale=scale, seed=bench.SEED)
model = bench.init_state(store).to(dev).eval()
tl = bench.batches(store, 1)[0]
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)


def time_prog(prog, reps=30, do_flush=True):
    s = [torch.cuda.Event(enable_timing=True) for _ in range(reps)]
    e = [torch.cuda.Event(enable_timing=True) for _ in range(reps)]
    for i in range(reps):
        if do_flush:
            flush.fill_(1.0)
        s[i].record(); prog.run(); e[i].record()
    torch.cuda.synchronize()
    return float(np.median([a.elapsed_time(b) for a, b in zip(s, e)])) * 1e3


for fuse in (False, True):
    model.runtime.fuse_scan = fuse
    res = model.encode(tl)
    torch.cuda.synchronize()
    ops = [o for o in res.program.ops if o.kind != lib.OP_H2D]
    full = lib.Program(); full.ops = ops
    print("fuse=%s rows=%d edges=%d ops=%d  forward: flush %.1f us, warm %.1f us" % (fuse, res.plan.R, res.plan.E, len(ops), time_prog(full), time_prog(full, do_flush=False)))
    for i, o in enumerate(ops):
        one = lib.Program(); one.ops = [o]
        kind = {1: "layer", 2: "gru", 8: "scan"}[o.kind]
        extra = ""
        if o.kind == 2:
            extra = "rows=%d" % (o.u.gru.row1 - o.u.gru.row0)
        print("   op %d %-5s %s flush %.1f us  warm %.1f us" % (i, kind, extra, time_prog(one), time_prog(one, do_flush=False)))
    if fuse:
        scan = [o for o in ops if o.kind == 8][0]
        n = scan.u.scan.n_steps
        for k in (1, 2, 4, n):
            scan.u.scan.n_steps = k
            one = lib.Program(); one.ops = [scan]
            print("   scan with %d steps: flush %.1f us warm %.1f us" % (k, time_prog(one), time_prog(one, do_flush=False)))
        scan.u.scan.n_steps = n
