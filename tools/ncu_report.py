"""Compact text summary of an .ncu-rep for profiles/: key raw metrics per kernel + opcode/stall
tables from the source page.   usage: ncu_report.py file.ncu-rep kernel_regex [kernel_regex ...]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "sm__cycles_active.avg"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
print(f"# ncu --set full --clock-control none : {rep}")
for r in rows[2:]:
    print("\n== " + r[hdr.index("Kernel Name")][:100])
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"{k:70s} {r[i]:>18s} {units[i]}")
for pat in sys.argv[2:]:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}"],
                         capture_output=True, text=True).stdout
    open("/tmp/_ncu_src.csv", "w").write(src)
    print(f"\n== source page summary, kernel regex {pat}")
    out = subprocess.run([sys.executable, __file__.replace("ncu_report.py", "ncu_src_summary.py"),
                          "/tmp/_ncu_src.csv", "12"], capture_output=True, text=True).stdout
    print(out)
