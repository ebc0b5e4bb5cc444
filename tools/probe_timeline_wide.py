"""Development probe (GPU): per-phase clock64 timeline of gru_scan_tcw_kernel (temp_b200/csrc/tc_wide.cu) on BASELINE config 3.
Builds a SEPARATE library with -DTEMP_TIMELINE (the product library carries no instrumentation) and prints, per step, the
median over CTAs of the cycles between consecutive phase marks of worker warp 0.

    python tools/probe_timeline_wide.py
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

from tests import test_gpu_fullsize as T

cfg = next(c for c in T.CONFIGS if c[0] == "config3_bigrrgcn_icews0515_nb100")
model, _, t_list = T._build(cfg)
res = model.encode(t_list)
torch.cuda.synchronize()
CT, W, ST, MK = 160, 8, 16, 10
buf = torch.zeros(CT * W * ST * MK, dtype=torch.int64, device="cuda")
L.temp_debug_timeline_wide.argtypes = [C.c_void_p]
PH = ["wait producers", "gather+convert+arrive", "gi/te loads issued", "wait MMA", "park+bar", "gates+store", "cta sync", "(next step starts)"]
for o in [o for o in res.program.ops if o.kind == lib.OP_GRU_SCAN]:
    one = lib.Program()
    one.ops = [o]
    for _ in range(3):
        one.run()
    buf.zero_()
    torch.cuda.synchronize()
    assert L.temp_debug_timeline_wide(C.c_void_p(buf.data_ptr())) == 0
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    one.run()
    e.record()
    torch.cuda.synchronize()
    L.temp_debug_timeline_wide(None)
    t = buf.view(CT, W, ST, MK).cpu().numpy().astype(np.float64)
    print("scan: %d steps, event time %.1f us" % (o.u.scan.n_steps, s.elapsed_time(e) * 1e3))
    for st in range(o.u.scan.n_steps):
        tw = t[:, 0, st, :]
        ok = (tw[:, 0] != 0) & (tw[:, 6] != 0)
        if ok.sum() == 0:
            continue
        tw = tw[ok]
        seg = [np.median(tw[:, k + 1] - tw[:, k]) for k in range(6)]
        nxt = t[:, 0, st + 1, 0][ok] if st + 1 < ST else np.zeros(ok.sum())
        gap = np.median((nxt - tw[:, 6])[nxt != 0]) if (nxt != 0).any() else float("nan")
        print("  step %2d (%3d CTAs): " % (st, int(ok.sum())) + "  ".join("%s %.0f" % (PH[k], seg[k]) for k in range(6)) +
              "  | end -> next step's top %.0f  | total %.0f" % (gap, np.median(tw[:, 6] - tw[:, 0])))
