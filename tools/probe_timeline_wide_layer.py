"""Development probe (GPU): per-phase clock64 timeline of rgcn_layer_tcw_kernel (temp_b200/csrc/tc_wide.cu) for every layer
launch of BASELINE config 3 (separate -DTEMP_TIMELINE library): median over CTAs of the cycles between the marks of worker
warps 0 (rows 0..31, feature quadrant 0) and 7.  (The buffer is indexed by blockIdx.x only: for a launch whose chain blocks
are spread over grid.y -- the Bi centre step -- the y blocks overwrite each other's marks and the line is meaningless.)

    python tools/probe_timeline_wide_layer.py
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
CT, W, ST, MK = 2200, 8, 16, 10
buf = torch.zeros(CT * W * ST * MK, dtype=torch.int64, device="cuda")
L.temp_debug_timeline_wide.argtypes = [C.c_void_p]
PH = ["prologue->pdl_wait", "stage operand", "ag loads issued", "wait GEMM1", "epilogue 1", "chain epilogues"]
for o in [o for o in res.program.ops if o.kind == lib.OP_LAYER]:
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
    print("layer rows %d chain_n %d: event time %.1f us (gather + tile kernel)" % (o.u.layer.row1 - o.u.layer.row0, o.u.layer.chain_n, s.elapsed_time(e) * 1e3))
    for w in (0, 7):
        tw = t[:, w, 0, :]
        ok = (tw[:, 0] != 0) & (tw[:, 6] != 0)
        tw = tw[ok]
        seg = [np.median(tw[:, k + 1] - tw[:, k]) for k in range(6)]
        print("   warp %d (%4d CTAs): " % (w, int(ok.sum())) + "  ".join("%s %.0f" % (PH[k], seg[k]) for k in range(6)) +
              "  | total %.0f cycles" % np.median(tw[:, 6] - tw[:, 0]))
