#!/usr/bin/env python3
"""ncu report -> per-kernel summary (JSON + CSV) of the counters the roofline record quotes.
   python tools/ncu_summary.py <report.ncu-rep> <out.json> [<out.csv>]
One entry per kernel name (last captured launch of each): duration, DRAM bytes per launch, shared-memory / L1 data pipe,
ALU / FMA (HSET2 runs on the FMA pipe) / FP64 pipes, issue slots, warps active, registers, bank conflicts, top stalls."""
import csv
import json
import subprocess
import sys

METRICS = {
    "gpu__time_duration.sum": "duration_ns",
    "dram__bytes_read.sum": "dram_bytes_read",
    "dram__bytes_write.sum": "dram_bytes_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed": "smem_l1_data_pipe_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "shared_wavefronts",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "shared_bank_conflicts",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "alu_pipe_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "fma_pipe_pct",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "fp64_pipe_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "lsu_pipe_pct",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "smsp__inst_executed.sum": "warp_instructions",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__occupancy_limit_shared_mem": "ctas_per_sm_limit_smem",
    "launch__occupancy_limit_registers": "ctas_per_sm_limit_regs",
    "launch__grid_size": "grid_size",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
}
SHORT = {"harris_box_kernel": "harris_box", "boxsum9_kernel": "boxsum_right", "select_corners_kernel": "select_corners", "describe_left_kernel": "describe_left",
         "stereo_match_kernel": "stereo_match", "stereo_match_split_kernel": "stereo_match",
         "stereo_match_binned_kernel": "stereo_match", "describe_left_binned_kernel": "describe_left", "bin_keypoints_kernel": "bin_keypoints"}


def main():
    rep, out_json = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    name_i = hdr.index("Kernel Name")
    scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}

    def val(r, metric):
        i = hdr.index(metric)
        return float(r[i].replace(",", "")) * scale.get(units[i], 1.0)
    res = {}
    for r in rows[2:]:
        if len(r) != len(hdr):
            continue
        kn = r[name_i]
        key = next((v for k, v in SHORT.items() if k in kn), kn.split("(")[0])
        d = {}
        for m, nice in METRICS.items():
            if m in hdr:
                try:
                    d[nice] = float(r[hdr.index(m)].replace(",", ""))
                except ValueError:
                    pass
        stalls = {h[len("smsp__average_warps_issue_stalled_"):].split("_per_issue_active")[0]: float(r[i].replace(",", ""))
                  for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and r[i] not in ("", "n/a")}
        d["top_stalls"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:4])
        d["dram_bytes_read"], d["dram_bytes_write"] = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
        d["dram_bytes_per_launch"] = d["dram_bytes_read"] + d["dram_bytes_write"]
        d["us_per_launch"] = val(r, "gpu__time_duration.sum")
        d.pop("duration_ns", None)
        res[key] = d
    json.dump(res, open(out_json, "w"), indent=1)
    if len(sys.argv) > 3:
        cols = ["kernel"] + sorted({k for d in res.values() for k in d if k != "top_stalls"})
        with open(sys.argv[3], "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(cols + ["top_stalls"])
            for k, d in res.items():
                w.writerow([k] + [d.get(c, "") for c in cols[1:]] + [json.dumps(d["top_stalls"])])
    print(json.dumps({k: {"us": round(v["us_per_launch"], 1), "dram_MB": round(v["dram_bytes_per_launch"] / 1e6, 1)} for k, v in res.items()}))


if __name__ == "__main__":
    main()
