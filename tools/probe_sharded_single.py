"""Development probe (one GPU): per-rank launch programs of a snapshot-sharded forward (GDELT-shaped config 5) timed alone --
how far phase 1 and the scan shrink with the rank's share of the rows (the exchanges are not run).

    python tools/probe_sharded_single.py [scale]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from temp_b200 import lib
from temp_b200.models import build_module
from temp_b200.snapshot import SnapshotStore
from temp_b200.sharding import make_shard_plan
dev = torch.device("cuda", 0)
scale = int(sys.argv[1]) if len(sys.argv) > 1 else 1
store = SnapshotStore.synthetic("gdelt", num_times=24 if scale == 1 else 18, scale=scale, seed=20201116 + 4)
a = bench.make_args(); a.train_seq_len = a.test_seq_len = 15
torch.manual_seed(123)
model = build_module(a, store.num_ents, store.num_rels, store.train).to(dev).eval()
t_list = [store.times[-3], store.times[-2]]
def timeit(prog, reps=30):
    for _ in range(5): prog.run()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps): prog.run()
    e.record(); torch.cuda.synchronize()
    return 1e3 * s.elapsed_time(e) / reps
single = model.encode(t_list)
ops = [o for o in single.replay.ops]
for i, o in enumerate(ops):
    p = lib.Program(); p.ops = [o]
    print("unsharded op %d kind %d: %.1f us" % (i, o.kind, timeit(p)))
plan = model.plan(t_list)
for world in (2, 8):
    shard = make_shard_plan(plan, world)
    for r in range(world if world == 2 else 2):
        res = model.runtime.build_sharded(model.plan(t_list), shard, r)
        p0 = lib.Program(); p0.ops = [o for o in res.programs[0].ops if o.kind != lib.OP_H2D]
        res.programs[0].run(); torch.cuda.synchronize()
        lo, hi = shard.rows_of(r)
        per_op = []
        for o in p0.ops:
            one = lib.Program()
            one.ops = [o]
            per_op.append(round(timeit(one), 1))
        print("world %d rank %d rows [%d,%d) of %d: phase-1 %.1f us (ops %s); scan %.1f us" % (world, r, lo, hi, plan.R, timeit(p0),
                                                                                              per_op, timeit(res.programs[1])))
