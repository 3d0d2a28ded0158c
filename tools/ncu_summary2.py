#!/usr/bin/env python
"""Per-kernel summary of an .ncu-rep with the pipe view that matters for this library (integer ALU pipe is half rate,
LSU wavefronts are one per cycle per SM):  python tools/ncu_summary2.py gpurun_out/prof.ncu-rep > profiles/<name>.txt"""
import csv, io, subprocess, sys

KEYS = [
    ("time_us", "gpu__time_duration.sum"),
    ("dram_rd_MB", "dram__bytes_read.sum"), ("dram_wr_MB", "dram__bytes_write.sum"),
    ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("issue_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("alu_pipe_pct", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
    ("fma_pipe_pct", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
    ("lsu_inst_pct", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
    ("lsu_wavefront_pct", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed"),
    ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("regs", "launch__registers_per_thread"), ("grid", "launch__grid_size"), ("block", "launch__block_size"),
    ("dyn_smem", "launch__shared_mem_per_block_dynamic"),
    ("warp_inst", "smsp__inst_executed.sum"),
    ("smem_ld_inst", "smsp__sass_inst_executed_op_shared_ld.sum"), ("smem_st_inst", "smsp__sass_inst_executed_op_shared_st.sum"),
    ("smem_ld_wavefronts", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum"),
    ("smem_st_wavefronts", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum"),
    ("smem_bank_conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
    ("branch_inst", "smsp__inst_executed_op_branch.sum"),
    ("sm_cycles", "sm__cycles_elapsed.avg"),
]
STALLS = ["barrier", "long_scoreboard", "short_scoreboard", "mio_throttle", "lg_throttle", "membar", "wait", "math_pipe_throttle",
          "no_instruction", "branch_resolving", "not_selected", "dispatch_stall", "sleeping", "tex_throttle", "drain", "imc_miss", "selected"]

def main():
    raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("==", r[idx["Kernel Name"]].split("(")[0])
        for label, key in KEYS:
            if key in idx:
                print(f"   {label:22s} {r[idx[key]]} {units[idx[key]]}")
        st = []
        for s in STALLS:
            key = f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio"
            if key in idx:
                st.append((float(r[idx[key]].replace(",", "")), s))
        st.sort(reverse=True)
        print("   stalls(warps per issue): " + ", ".join(f"{s}={v:.2f}" for v, s in st[:7]))

if __name__ == "__main__":
    main()
