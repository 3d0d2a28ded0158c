#!/usr/bin/env python
"""Compact per-kernel summary of an .ncu-rep (run where ncu is installed; no GPU needed):
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--stalls] > profiles/<name>.txt
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("time_us", "gpu__time_duration.sum"),
    ("dram_rd_MB", "dram__bytes_read.sum"),
    ("dram_wr_MB", "dram__bytes_write.sum"),
    ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("sm_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("issue_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("regs", "launch__registers_per_thread"),
    ("occ_lim_regs", "launch__occupancy_limit_registers"),
    ("occ_lim_smem", "launch__occupancy_limit_shared_mem"),
    ("grid", "launch__grid_size"),
    ("block", "launch__block_size"),
    ("warp_inst", "smsp__inst_executed.sum"),
    ("smem_bank_conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
    ("l2_hit_pct", "lts__t_sector_hit_rate.pct"),
]
STALLS = ["barrier", "long_scoreboard", "short_scoreboard", "mio_throttle", "lg_throttle", "membar", "wait",
          "math_pipe_throttle", "no_instruction", "branch_resolving", "not_selected", "dispatch_stall", "sleeping",
          "tex_throttle", "drain", "imc_miss", "selected"]


def to_bytes(v, unit):
    f = float(v.replace(",", ""))
    u = unit.lower()
    return f * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0]
        print(f"== {name}")
        vals = {}
        for label, key in KEYS:
            if key in idx:
                v, u = r[idx[key]], units[idx[key]]
                if label in ("dram_rd_MB", "dram_wr_MB"):
                    vals[label] = to_bytes(v, u) / 1e6
                    print(f"   {label:22s} {vals[label]:.2f}")
                elif label == "time_us":
                    f = float(v.replace(",", "")) * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)
                    vals[label] = f
                    print(f"   {label:22s} {f:.2f}")
                else:
                    print(f"   {label:22s} {v} {u}")
        if "time_us" in vals and "dram_rd_MB" in vals:
            tot = vals["dram_rd_MB"] + vals["dram_wr_MB"]
            print(f"   {'dram_total_MB':22s} {tot:.2f}   ({tot / vals['time_us'] * 1e3:.0f} GB/s under ncu)")
        st = []
        for s in STALLS:
            key = f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio"
            if key in idx:
                st.append((float(r[idx[key]].replace(",", "")), s))
        st.sort(reverse=True)
        print("   stalls(warps per issue): " + ", ".join(f"{s}={v:.2f}" for v, s in st[:7]))


if __name__ == "__main__":
    main()
