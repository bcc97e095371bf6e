"""Summarise an ncu report (one row per profiled launch) as CSV on stdout:
   python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/xxx.csv"""
import csv, io, subprocess, sys
WANT = [('Kernel Name', 'kernel'), ('launch__grid_size', 'grid'), ('launch__block_size', 'block'), ('launch__registers_per_thread', 'regs'),
        ('launch__shared_mem_per_block_dynamic', 'dyn_smem'), ('launch__shared_mem_per_block_static', 'static_smem'),
        ('gpu__time_duration.sum', 'time'), ('dram__bytes_read.sum', 'dram_read'), ('dram__bytes_write.sum', 'dram_write'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram_pct'), ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l2_pct'),
        ('l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1_pct'), ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm_pct'),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue_pct'), ('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'alu_pct'),
        ('sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'fma_pct'), ('sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'lsu_pct'),
        ('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'xu_pct'), ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps_active_pct'),
        ('smsp__inst_executed.sum', 'warp_inst'), ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor_pct')]
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h, units = rows[0], rows[1]
w = csv.writer(sys.stdout)
cols = [(h.index(k), n, units[h.index(k)]) for k, n in WANT if k in h]
w.writerow(['%s[%s]' % (n, u) if u else n for _, n, u in cols])
for r in rows[2:]:
    w.writerow([r[i].split('(')[0] if n == 'kernel' else r[i] for i, n, _ in cols])
