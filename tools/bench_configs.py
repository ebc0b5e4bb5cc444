"""Secondary measurement: region R1 (evaluate_embed) of every BASELINE.json configuration at its full size on synthetic
snapshots of the reference's dataset shapes -- device time of the CUDA forward (CUDA events, L2 flushed per run, median of
30) next to the oracle port of the reference's CPU path on the host cores.  One JSON line per configuration."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from tests import test_gpu_fullsize as T

flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
torch.set_num_threads(os.cpu_count() or 1)
for cfg in T.CONFIGS[:6]:
    model, oracle, t_list = T._build(cfg)
    res = model.encode(t_list)
    torch.cuda.synchronize()
    prog = res.program
    ts = []
    for _ in range(30):
        flush.fill_(1.0)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        prog.run()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ms = float(np.median(ts))
    with torch.no_grad():
        oracle.evaluate_embed(t_list)
        t0 = time.perf_counter()
        n = 0
        while time.perf_counter() - t0 < 2.0:
            oracle.evaluate_embed(t_list)
            n += 1
        cpu_ms = 1e3 * (time.perf_counter() - t0) / n
    print(json.dumps({"config": cfg[0], "module": cfg[1], "shape": cfg[2], "D": cfg[3], "n_bases": cfg[4], "seq_len": cfg[5],
                      "batch": len(t_list), "rows": int(res.plan.R), "edges": int(res.plan.E),
                      "kernel_launches": prog.kernel_count(), "gpu_ms": ms, "gpu_edges_per_s": res.plan.E / (ms * 1e-3),
                      "cpu_oracle_ms": cpu_ms, "cpu_edges_per_s": res.plan.E / (cpu_ms * 1e-3), "cpu_cores": os.cpu_count(),
                      "path": "tcgen05" if cfg[3] == 128 else ("fp32 SIMT (TEMP_WIDE_TC=0)" if os.environ.get("TEMP_WIDE_TC") == "0"
                                                                else "tcgen05, 64-row tiles (tc_wide.cu)")}))
