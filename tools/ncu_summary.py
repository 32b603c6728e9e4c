"""Condense an .ncu-rep (read here, no GPU) into the handful of numbers DESIGN.md / profiles/ quote.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [out.txt]"""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
KEEP = ['Kernel Name', 'Block Size', 'Grid Size', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
        'smsp__warps_active.avg.per_cycle_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sass__inst_executed_shared_loads', 'sass__inst_executed_shared_stores',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active']
out = []
for vals in rows[2:]:
    for h, u, v in zip(hdr, units, vals):
        if h in KEEP or h.startswith('smsp__average_warps_issue_stalled'):
            out.append("%-92s %-16s %s" % (h, u, v))
    out.append("")
txt = "\n".join(out)
if len(sys.argv) > 2:
    open(sys.argv[2], "a").write(txt + "\n")
print(txt)
