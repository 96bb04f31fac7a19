"""`ncu -i X.ncu-rep --page raw --csv` -> the per-launch summary JSON kept under profiles/ (same keys as
profiles/r02_ncu_full_kernels.json).  usage: python tools/ncu_summary.py raw.csv "command" out.json"""
import csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
KEYS = {
    "duration_us": "gpu__time_duration.sum", "dram_read_MB": "dram__bytes_read.sum", "dram_write_MB": "dram__bytes_write.sum",
    "tensor_pipe_pct_active": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "regs": "launch__registers_per_thread", "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram_throughput_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm_throughput_pct": "sm__throughput.avg.pct_of_peak_sustained_elapsed", "warp_instructions": "sm__inst_executed.sum",
    "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smem_bank_conflicts": "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l2_read_sectors_from_sm": "lts__t_sectors_srcunit_tex_op_read.sum", "l2_hit_rate_pct": "lts__t_sector_hit_rate.pct",
    "dyn_smem_KB": "launch__shared_mem_per_block_dynamic", "sm_cycles_active": "sm__cycles_active.avg",
    "occupancy_limit_registers": "launch__occupancy_limit_registers", "occupancy_limit_shared_mem": "launch__occupancy_limit_shared_mem",
    "stall_long_scoreboard": "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "stall_short_scoreboard": "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "stall_wait": "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "stall_barrier": "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
}
def num(v):
    try:
        return float(v.replace(',', ''))
    except ValueError:
        return v
def scale(v, unit, want):
    f = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit) if want == "MB" else \
        {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}.get(unit) if want == "us" else \
        {"byte": 1e-3, "Kbyte": 1.0, "Mbyte": 1e3}.get(unit) if want == "KB" else None
    return v * f if (f is not None and isinstance(v, float)) else v
out = {"command": sys.argv[2], "launches": []}
for r in rows[2:]:
    d, u = dict(zip(hdr, r)), dict(zip(hdr, units))
    e = {"kernel": d.get("Kernel Name"), "grid": d.get("Grid Size"), "block": d.get("Block Size")}
    for k, m in KEYS.items():
        if m in d:
            v = num(d[m])
            want = "MB" if k.endswith("_MB") else "us" if k.endswith("_us") else "KB" if k.endswith("_KB") else None
            e[k] = round(scale(v, u[m], want), 3) if isinstance(scale(v, u[m], want), float) else v
    out["launches"].append(e)
json.dump(out, open(sys.argv[3], "w"), indent=1)
for e in out["launches"]:
    print(e["kernel"][:60], e.get("duration_us"), "us  dram r/w MB", e.get("dram_read_MB"), e.get("dram_write_MB"), "regs", e.get("regs"),
          "warps %", e.get("warps_active_pct"), "issue %", e.get("issue_active_pct"))
