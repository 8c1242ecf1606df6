#!/usr/bin/env python3
"""Turn an .ncu-rep brought back from the GPU box into the small tracked summaries under profiles/.

    python profiles/summarize_ncu.py gpurun_out/prof_x.ncu-rep profiles/r01_x   [--traffic-key pga_sandwich]

Writes <out>.raw.csv (the full `--page raw` table of the captured launches), <out>.summary.json (the
metrics DESIGN.md / bench.py quote) and, with --traffic-key, records the measured DRAM traffic per
launch in profiles/traffic.json, which bench.py reports as roofline.traffic.
"""
import csv
import io
import json
import os
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg.per_second",
    "l1tex__t_bytes.sum", "lts__t_bytes.sum", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]
TO_BYTES = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    key = sys.argv[sys.argv.index("--traffic-key") + 1] if "--traffic-key" in sys.argv else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    raw = raw[raw.index('"ID"'):]
    with open(out + ".raw.csv", "w") as f:
        f.write(raw)
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    launches = []
    for r in body:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                try:
                    d[k] = {"value": float(r[i].replace(",", "")), "unit": units[i]}
                except ValueError:
                    d[k] = {"value": r[i], "unit": units[i]}
        launches.append(d)
    summary = {"report": os.path.basename(rep), "launches": launches}
    if launches and "dram__bytes_read.sum" in launches[0]:
        def b(d, k):
            return d[k]["value"] * TO_BYTES.get(d[k]["unit"], 1)
        per = [b(d, "dram__bytes_read.sum") + b(d, "dram__bytes_write.sum") for d in launches]
        summary["dram_traffic_bytes_per_launch"] = sum(per) / len(per)
        if key:
            path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "traffic.json")
            t = json.load(open(path)) if os.path.exists(path) else {}
            t[key] = summary["dram_traffic_bytes_per_launch"]
            json.dump(t, open(path, "w"), indent=1, sort_keys=True)
    json.dump(summary, open(out + ".summary.json", "w"), indent=1)
    print(json.dumps(summary, indent=1)[:1500])


if __name__ == "__main__":
    main()
