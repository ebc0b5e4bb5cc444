"""Development probe (GPU): per-phase clock64 timeline of the tcgen05 kernels on the bench workload.

Builds a SEPARATE library with -DTEMP_TIMELINE (the product library carries no instrumentation), runs each op
of the forward alone and prints, per kernel, the median over CTAs of the cycle count at which each phase mark
was reached (relative to the CTA's first mark), for worker warp 0 and for the MMA-issuing warp.

    python tools/probe_timeline.py [scale]
"""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from temp_b200 import build as B
from temp_b200 import lib

TL_LIB = os.path.join(ROOT, "tools", "libtemp_b200_tl.so")
cmd = [B.nvcc_path(), "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-DTEMP_TIMELINE",
       "-Xcompiler", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"), "-o", TL_LIB] + B.SOURCES
if not os.path.exists(TL_LIB) or any(os.path.getmtime(p) > os.path.getmtime(TL_LIB) for p in B.SOURCES + B.HEADERS):
    subprocess.run(cmd, check=True)
L = lib.load(TL_LIB)

import bench
from temp_b200.snapshot import SnapshotStore

dev = torch.device("cuda", 0)
scale = int(sys.argv[1]) if len(sys.argv) > 1 else 1
store = SnapshotStore.synthetic("icews14", num_times=40 if scale == 1 else 16, scale=scale, seed=bench.SEED)
model = bench.init_state(store).to(dev).eval()
tl = bench.batches(store, 1)[0]
res = model.encode(tl)
torch.cuda.synchronize()
WARPS, SLOTS = 20, 64  # kTlWarps, kTlSlots of the probe build
MAXCTA = 4096
buf = torch.zeros(MAXCTA * WARPS * SLOTS, dtype=torch.int64, device=dev)
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
L.temp_debug_timeline.argtypes = [C.c_void_p]
ops = [o for o in res.program.ops if o.kind != lib.OP_H2D]
names = {1: "layer", 2: "gru", 8: "scan"}
for cold in (True, False):
    for i, o in enumerate(ops):
        one = lib.Program()
        one.ops = [o]
        for _ in range(3):
            one.run()
        buf.zero_()
        if cold:
            flush.fill_(1.0)
        torch.cuda.synchronize()
        assert L.temp_debug_timeline(C.c_void_p(buf.data_ptr())) == 0
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        one.run()
        e.record()
        torch.cuda.synchronize()
        L.temp_debug_timeline(None)
        t = buf.view(MAXCTA, WARPS, SLOTS).cpu().numpy()
        used = t[:, 0, 0] != 0
        n = int(used.sum())
        print("== op %d %s  %s  ctas=%d  event time %.1f us" % (i, names.get(o.kind, o.kind), "cold-L2" if cold else "warm", n,
                                                               s.elapsed_time(e) * 1e3))
        if n == 0:
            continue
        t = t[used]
        g0 = t[:, :, 63].astype(np.float64)
        g0 = g0[g0 > 0]
        print("   CTA start spread (globaltimer): %.2f us" % ((g0.max() - g0.min()) / 1e3))
        for w, label in ((0, "worker warp 0"), (7, "worker warp 7"), (9, "mma warp (layer)"), (8, "warp 8"), (15, "scan control warp")):
            rel = t[:, w, :63].astype(np.float64) - t[:, w, 0:1].astype(np.float64)
            ok = t[:, w, :63] != 0
            line = []
            for sl in range(63):
                if ok[:, sl].sum() > 0:
                    v = rel[ok[:, sl], sl]
                    line.append("%d:%.0f/%.0f" % (sl, np.median(v), v.max()))
            if line:
                print("   %-18s (slot:median/max cycles) %s" % (label, " ".join(line)))
