"""Reads an .ncu-rep (ncu -i … --page raw --csv) and prints, per kernel launch, the figures profiles/ keeps:
duration, DRAM bytes, tensor-pipe / issue utilisation, shared-memory bank conflicts and the top stall reasons.

    python tools/ncu_summarise.py gpurun_out/r2_full_x1.ncu-rep [--json out.json]
"""
import csv
import io
import json
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("lts__t_bytes.sum", "l2_bytes"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed", "mem_pct"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor_inst"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pct"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor_hmma_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__cluster_size", "cluster"),
    ("smsp__cycles_active.avg", "cycles_active"),
]


def to_float(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return None


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, units, body = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(head)}
    stall_cols = [h for h in head if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")
                  or h.startswith("smsp__average_warp_latency_issue_stalled_") and h.endswith(".ratio")]
    res = []
    for r in body:
        d = {"kernel": r[col["Kernel Name"]].split("(")[0], "id": r[col["ID"]]}
        for m, k in WANT:
            if m in col:
                v = to_float(r[col[m]])
                d[k] = v
                d[k + "_unit"] = units[col[m]]
        st = sorted(((to_float(r[col[h]]) or 0.0, h) for h in stall_cols), reverse=True)[:5]
        d["top_stalls"] = [(h.replace("smsp__average_warps_issue_stalled_", "").replace("smsp__average_warp_latency_issue_stalled_", "")
                            .replace("_per_issue_active.ratio", "").replace(".ratio", ""), round(v, 2)) for v, h in st]
        res.append(d)
    for d in res:
        print(json.dumps(d))
    if "--json" in sys.argv:
        json.dump(res, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)


if __name__ == "__main__":
    main()
