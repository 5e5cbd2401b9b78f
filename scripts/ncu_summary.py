"""print the key ncu metrics (and top stall reasons) of every kernel in a .ncu-rep"""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__cluster_size',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'sm__maximum_warps_per_active_cycle_pct',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ldgsts.sum', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
        'l1tex__m_xbar2l1tex_read_bytes.sum', 'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum']
for r in rows[2:]:
    print("==", r[hdr.index('Kernel Name')][:90])
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f"   {w:70s} {r[i]:>16s} {units[i]}")
    stalls = []
    for i, h in enumerate(hdr):
        if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio') or (
                h.startswith('smsp__average_warp_latency_issue_stalled') and h.endswith('.ratio')):
            try:
                stalls.append((float(r[i]), h))
            except ValueError:
                pass
    for v, h in sorted(stalls, reverse=True)[:8]:
        print(f"   stall {v:8.2f}  {h.replace('smsp__average_warps_issue_stalled_', '').replace('smsp__average_warp_latency_issue_stalled_', '')}")
