"""Print the metrics we track from an ncu report: python scripts/ncu_keys.py <report.ncu-rep> [kernel-row]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
row = int(sys.argv[2]) if len(sys.argv) > 2 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2 + row]
KEYS = """gpu__time_duration.sum dram__bytes_read.sum dram__bytes_write.sum smsp__inst_executed.sum
smsp__issue_active.avg.pct_of_peak_sustained_active sm__warps_active.avg.pct_of_peak_sustained_active
launch__registers_per_thread l1tex__data_pipe_lsu_wavefronts_mem_shared.sum
l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed
sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active
gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed sm__throughput.avg.pct_of_peak_sustained_elapsed
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active lts__throughput.avg.pct_of_peak_sustained_elapsed launch__grid_size launch__block_size sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active""".split()
for i, h in enumerate(hdr):
    if h in KEYS or ("pcsamp_warps_issue_stalled" in h and "not_issued" not in h):
        print(f"{h:86s} {vals[i]:>18s} {units[i]}")
