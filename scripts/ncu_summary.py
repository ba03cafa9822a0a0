"""Extracts the judged numbers from an .ncu-rep (run here, no GPU needed) into a text summary.
usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.txt"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__cycles_elapsed.max", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
STALL = "smsp__average_warps_issue_stalled_"

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("kernel:", d.get("Kernel Name", "?")[:150])
    for k in WANT:
        if k in d:
            print("  %-72s %-10s %s" % (k, units[hdr.index(k)], d[k]))
    stalls = sorted(((float(d[h]), h[len(STALL):-len("_per_issue_active.ratio")]) for h in hdr
                     if h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and d[h] not in ("", "n/a")), reverse=True)
    print("  warp stalls per issue (top):", ", ".join("%s %.2f" % (n, v) for v, n in stalls[:7]))
    rd, wr = d.get("dram__bytes_read.sum"), d.get("dram__bytes_write.sum")
    if rd and wr:
        ur, uw = units[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_write.sum")]
        sc = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        print("  traffic (dram read+write) bytes: %.0f" % (float(rd) * sc[ur] + float(wr) * sc[uw]))
