"""Development probe (GPU): rgcn_layer_tcw_kernel alone (no aggregation) at d = 200 with a 600-wide chained GEMM, by number of
64-row tiles -- does the time per CTA depend on how many CTAs stream weight chunks at once (chip-level L2 -> SM throughput)
or not (per-SM ring depth / latency)?  One JSON line per tile count."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from temp_b200 import lib

L = lib.load()
D, N = 200, int(sys.argv[1]) if len(sys.argv) > 1 else 600
dev = "cuda"
g = torch.Generator(device="cpu").manual_seed(1)
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)


def pack(w):
    k, n = int(w.shape[0]), int(w.shape[1])
    buf = torch.empty(L.temp_packed_weights_bytes(k, n), dtype=torch.uint8, device=dev)
    lib.check(L.temp_pack_weights(C.c_void_p(w.data_ptr()), k, n, C.c_void_p(buf.data_ptr()), C.c_void_p(lib.current_stream())), "pack")
    return buf


loop_w = (torch.randn(D, D, generator=g) / D ** 0.5).to(dev)
chain_w = (torch.randn(D, N, generator=g) / D ** 0.5).to(dev) if N else None
pk1, pk2 = pack(loop_w), (pack(chain_w) if N else None)
for tiles in (4, 16, 37, 74, 111, 148, 296):
    R = 64 * tiles
    x = torch.randn(R, D, generator=g).to(dev)
    h = torch.empty(R, D, device=dev)
    ch = torch.empty(R, max(N, 4), device=dev)
    a = lib.RgcnLayerArgs()
    a.row0, a.row1, a.d = 0, R, D
    a.n_terms = 1
    a.terms[0].a, a.terms[0].w, a.terms[0].w_packed = x.data_ptr(), loop_w.data_ptr(), pk1.data_ptr()
    a.activation = lib.ACT_RELU
    a.h_out = h.data_ptr()
    if N:
        a.chain_w, a.chain_out, a.chain_n, a.chain_ld, a.chain_w_packed = chain_w.data_ptr(), ch.data_ptr(), N, N, pk2.data_ptr()
    ts = []
    for _ in range(12):
        flush.fill_(1.0)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        lib.check(L.temp_rgcn_layer_fwd(C.byref(a), C.c_void_p(lib.current_stream())), "layer")
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    n_mb = (N + 127) // 128
    gy = max(1, min(n_mb, 148 // tiles)) if N else 1
    chunks = 14 + 7 * ((n_mb + gy - 1) // gy)
    print(json.dumps({"tiles": tiles, "grid_y": gy, "ctas": tiles * gy, "chunks_per_cta": chunks, "us": float(np.median(ts)),
                      "us_per_chunk": float(np.median(ts)) / chunks / max(1, (tiles * gy + 147) // 148)}))
