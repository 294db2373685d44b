"""Condense an .ncu-rep (ncu --set full) into the handful of numbers DESIGN.md / bench.py quote.
Usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep [label] > profiles/rNN_x.txt"""
import csv, subprocess, sys, io
rep = sys.argv[1]
label = sys.argv[2] if len(sys.argv) > 2 else rep
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct", "smsp__warps_eligible.avg.per_cycle_active"]
print(f"# ncu --set full summary: {label}")
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    print(f"kernel: {name[:160]}")
    for k in keys:
        if k in hdr:
            i = hdr.index(k)
            print(f"  {k:75s} {r[i]:>16s} {units[i]}")
    print("  warp stall reasons (average warps per issue-active cycle):")
    st = [(h.split("issue_stalled_")[1].split("_per_issue")[0], float(r[i])) for i, h in enumerate(hdr)
          if "average_warps_issue_stalled" in h and "not_issued" not in h]
    for n, v in sorted(st, key=lambda t: -t[1]):
        if v >= 0.05:
            print(f"    {n:28s} {v:6.2f}")
