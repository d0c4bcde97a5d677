"""ncu report -> profiles/<tag>_ncu_summary.txt + profiles/<tag>_traffic.json (per-launch DRAM traffic of each captured kernel).
usage: python scripts/ncu_summary.py gpurun_out/prof_r1c.ncu-rep r1c"""
import csv, io, json, subprocess, sys
rep, tag = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
M = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
     "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
     "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__grid_size", "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
     "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
     "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
     "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]
def num(x, u):
    v = float(x.replace(",", ""))
    return v * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ms": 1e3, "us": 1.0, "ns": 1e-3}.get(u, 1.0)
out, traffic = [], {}
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")].split("(")[0]
    out.append(f"== {name}")
    vals = {}
    for m in M:
        if m in hdr:
            i = hdr.index(m); vals[m] = num(r[i], units[i]); out.append(f"   {m:95s} {r[i]} {units[i]}")
    t_us = vals.get("gpu__time_duration.sum", 0.0); rd, wr = vals.get("dram__bytes_read.sum", 0.0), vals.get("dram__bytes_write.sum", 0.0)
    if t_us: out.append(f"   -> DRAM traffic {(rd + wr) / 1e6:.1f} MB per launch = {(rd + wr) / t_us / 1e3:.0f} GB/s under ncu (cold caches, serialised)")
    traffic[name] = {"dram_bytes_read": rd, "dram_bytes_write": wr, "duration_us_under_ncu": t_us}
open(f"profiles/{tag}_ncu_summary.txt", "w").write("\n".join(out) + "\n")
json.dump(traffic, open(f"profiles/{tag}_traffic.json", "w"), indent=1)
print("\n".join(out))
