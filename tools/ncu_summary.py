"""Text summary of one kernel of an .ncu-rep (the files under profiles/ are made with it):
   python tools/ncu_summary.py report.ncu-rep > profiles/rNN_ncu_<kernel>_summary.txt"""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, u, v = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_config_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum"]
for k in want:
    for i, name in enumerate(h):
        if name == k:
            print("%-72s %20s %s" % (k, v[i], u[i]))


def f(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return 0.0


st = [(h[i], v[i]) for i in range(len(h)) if "pcsamp_warps_issue_stalled" in h[i] and not h[i].endswith("not_issued")]
tot = sum(f(x[1]) for x in st) or 1.0
print("warp stall samples (smsp__pcsamp_warps_issue_stalled_*), share of all samples:")
for n, val in sorted(st, key=lambda t: -f(t[1]))[:10]:
    print("  %-40s %10s  %5.1f %%" % (n.replace("smsp__pcsamp_warps_issue_stalled_", ""), val, 100 * f(val) / tot))
