#!/usr/bin/env python
"""Summarise an ncu report (one `--set full` capture per kernel) into the JSON files bench.py and the judge read.

    python scripts/ncu_summary.py REPORT.ncu-rep OUT_SUMMARY.json [TRAFFIC.json]
"""
import csv, json, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
traffic_out = sys.argv[3] if len(sys.argv) > 3 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_alu.sum",
    "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_xu.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
]
summary, traffic = {"report": rep, "command": "ncu --set full --clock-control none --import-source on", "kernels": []}, {}
KEY = {"tnf_forward": "forward", "tnf_backward_prop": "backward_prop", "tnf_backward_field": "backward_field",
       "tnf_wgrad": "wgrad", "tnf_adam": "adam", "tnf_losses": "losses", "tnf_rays": "raygen", "tnf_post": "post"}
for r in rows[2:]:
    name = r[idx["Kernel Name"]]
    k = {"kernel": name.split("(")[0], "metrics": {}}
    for w in WANT:
        if w in idx and r[idx[w]] != "":
            try:
                v = float(r[idx[w]].replace(",", ""))
            except ValueError:
                v = r[idx[w]]
            k["metrics"][w] = {"value": v, "unit": units[idx[w]]}
    summary["kernels"].append(k)
    m = k["metrics"]

    def to_bytes(e):
        u = e["unit"].lower()
        f = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
        return e["value"] * f

    if "dram__bytes_read.sum" in m:
        for pat, key in KEY.items():
            if pat in name and key not in traffic:
                traffic[key] = {"kernel": k["kernel"],
                                "dram_bytes_per_launch": to_bytes(m["dram__bytes_read.sum"]) + to_bytes(m["dram__bytes_write.sum"]),
                                "gpu_time_ms_under_ncu": m["gpu__time_duration.sum"]["value"] / (1e6 if m["gpu__time_duration.sum"]["unit"] == "ns" else 1e3 if m["gpu__time_duration.sum"]["unit"] == "us" else 1)}
json.dump(summary, open(out, "w"), indent=1)
if traffic_out:
    json.dump(traffic, open(traffic_out, "w"), indent=1)
for k in summary["kernels"]:
    m = k["metrics"]
    g = lambda n: m.get(n, {}).get("value")
    print(f"{k['kernel'][:44]:44s} t={g('gpu__time_duration.sum')} {m.get('gpu__time_duration.sum',{}).get('unit')} regs={g('launch__registers_per_thread')} "
          f"warps%={g('sm__warps_active.avg.pct_of_peak_sustained_active')} issue%={g('smsp__issue_active.avg.pct_of_peak_sustained_active')} "
          f"inst={g('smsp__inst_executed.sum')} l1hit={g('l1tex__t_sector_hit_rate.pct')} tensor%={g('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active')}")
