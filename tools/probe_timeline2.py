"""Development probe (GPU): per-phase clock64 timeline of gru_scan_tm_kernel (temp_b200/csrc/tc_scan2.cu) on the bench
workload.  Builds a SEPARATE library with -DTEMP_TIMELINE (the product library carries no instrumentation) and prints, per
warp of interest, the median over CTAs of the cycles between consecutive phase marks of tile steps 2..6 of a pipeline.

    python tools/probe_timeline2.py [scale]
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
print("rows %d  partitions %d  tile %d" % (res.plan.R, res.plan.scan_parts.shape[0], res.plan.scan_tile))
WARPS, SLOTS = 16, 64
MAXCTA = 1024
buf = torch.zeros(MAXCTA * WARPS * SLOTS, dtype=torch.int64, device=dev)
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
L.temp_debug_timeline2.argtypes = [C.c_void_p]
PH = ["top", "gi-req+stage-wait", "convert+bar", "mma-issue", "h0+mma-wait", "tmem->ex+bar", "gates+st+send"]
NPH, STRIDE, NST = 7, 8, 7
for o in [o for o in res.program.ops if o.kind == lib.OP_GRU_SCAN]:
    one = lib.Program()
    one.ops = [o]
    for cold in (True, False):
        for _ in range(3):
            one.run()
        buf.zero_()
        if cold:
            flush.fill_(1.0)
        torch.cuda.synchronize()
        assert L.temp_debug_timeline2(C.c_void_p(buf.data_ptr())) == 0
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        one.run()
        e.record()
        torch.cuda.synchronize()
        L.temp_debug_timeline2(None)
        t = buf.view(MAXCTA, WARPS, SLOTS).cpu().numpy()
        used = t[:, 0, 0] != 0
        print("== scan %s  ctas=%d  event time %.1f us" % ("cold-L2" if cold else "warm", int(used.sum()), s.elapsed_time(e) * 1e3))
        t = t[used].astype(np.float64)
        for w in (0, 1, 3, 4, 8, 12):
            tw = t[:, w, :]
            ok = tw[:, 3 + STRIDE * 1] != 0           # CTAs whose pipeline reached tile step 3
            if ok.sum() == 0:
                print("   warp %2d: pipeline idle" % w)
                continue
            tw = tw[ok]
            print("   warp %2d (pipe %d, %d CTAs): prologue cluster-sync %.0f, W->TMEM+indices %.0f" %
                  (w, w // 4, int(ok.sum()), np.median(tw[:, 1] - tw[:, 0]), np.median(tw[:, 2] - tw[:, 1])))
            for st in range(NST):
                base = 3 + STRIDE * st
                if (tw[:, base] == 0).any() or (tw[:, base + NPH - 1] == 0).any():
                    continue
                seg = []
                for k in range(1, NPH):
                    a, b = tw[:, base + k - 1], tw[:, base + k]
                    good = (a != 0) & (b != 0)
                    seg.append("%s %.0f" % (PH[k], np.median((b - a)[good])) if good.any() else "%s -" % PH[k])
                tot = np.median(tw[:, base + NPH - 1] - tw[:, base])
                if st + 1 < NST and (tw[:, base + STRIDE] != 0).all():
                    tot = np.median(tw[:, base + STRIDE] - tw[:, base])        # top to top: the whole tile step
                print("      tile step %d: total %.0f | %s" % (st + 2, tot, " | ".join(seg)))
