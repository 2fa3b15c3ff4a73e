"""Print the metrics that matter from an .ncu-rep (ncu --page raw --csv), one column per captured kernel launch."""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
WANT = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__occupancy_limit_warps",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_alu.sum",
        "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_uniform.sum", "sm__inst_executed_pipe_cbu.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
WANT += sorted(h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"))
for w in WANT:
    if w not in idx:
        continue
    vals = [r[idx[w]] for r in data]
    if w == "Kernel Name":
        vals = [v.replace("void cmh::<unnamed>::", "")[:34] for v in vals]
    else:
        try:
            vals = ["%.4g" % float(v.replace(",", "")) for v in vals]
        except ValueError:
            pass
    name = w.replace("smsp__average_warps_issue_stalled_", "stall.").replace("_per_issue_active.ratio", "")
    if name.startswith("stall.") and all(float(v) < 0.05 for v in vals):
        continue
    print("%-58s %-8s %s" % (name[:58], units[idx[w]][:8], "  ".join("%-12s" % v for v in vals)))
