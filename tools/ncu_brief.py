"""Short per-kernel digest of an .ncu-rep: python tools/ncu_brief.py rep [--cols 0,1]  (launch indices to keep)"""
import csv, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
if '--cols' in sys.argv:
    keep = [int(c) for c in sys.argv[sys.argv.index('--cols') + 1].split(',')]
    rows = rows[:2] + [rows[2 + c] for c in keep]
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__shared_mem_per_block_dynamic', 'launch__shared_mem_per_block_static',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sectors.sum', 'lts__t_sector_hit_rate.pct',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum',
        'smsp__average_warp_latency_per_inst_issued.ratio', 'sm__cycles_elapsed.max']
for i, h in enumerate(hdr):
    if h in want:
        print(f"{h}: {[r[i] for r in rows[2:]]}")
st = {}
for i, h in enumerate(hdr):
    if 'issue_stalled' in h and h.endswith('per_issue_active.ratio'):
        st[h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')] = [float(r[i]) for r in rows[2:]]
for k in range(len(rows) - 2):
    print(f"stalls[{k}]:", ", ".join(f"{n} {v[k]:.2f}" for n, v in sorted(st.items(), key=lambda x: -x[1][k])[:8]))
