"""Print the metrics we care about from an .ncu-rep (first kernel): python tools/ncu_summary.py rep [out.txt]"""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
KEYS = ['Kernel Name', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.avg.per_second', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'sm__inst_executed.avg.per_cycle_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'lts__t_sectors_op_red.sum', 'lts__t_sectors_op_atom.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_red.sum',
        'lts__t_bytes.sum']
out = []
for row in rows[2:3]:
    d = dict(zip(hdr, row))
    u = dict(zip(hdr, units))
    for k in KEYS:
        if k in d:
            out.append(f"{k}: {d[k]} {u[k]}".rstrip())
    for k in hdr:
        if 'issue_stalled' in k and k.endswith('per_issue_active.ratio'):
            try:
                if float(d[k]) >= 0.3:
                    out.append(f"{k}: {d[k]}")
            except ValueError:
                pass
text = "\n".join(out)
print(text)
if len(sys.argv) > 2:
    open(sys.argv[2], "a").write(text + "\n")
