"""Selected per-launch metrics of an `ncu --set full` report, as text (the .ncu-rep itself stays in gpurun_out/).

usage: python scripts/ncu_metrics.py gpurun_out/prof_r01_conv_tc.ncu-rep > profiles/r01_ncu_conv_tc.txt
"""
import csv, io, subprocess, sys

KEEP = [
    'ID', 'Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum',
    'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_tensor.sum',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
    'l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum',
    'l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum.per_second',
    'l1tex__t_sector_hit_rate.pct', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
    'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
    'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
    'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
    'smsp__warps_active.avg.per_cycle_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'smsp__cycles_active.avg', 'smsp__inst_executed.sum',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    print(f'# ncu --set full --clock-control none  (gpurun, B200) -- selected metrics per launch; source report: {rep}')
    for name in KEEP:
        if name not in hdr:
            continue
        i = hdr.index(name)
        vals = [r[i].replace('void raft::', '').replace('raft::', '') for r in data]
        print(f'{name:<95} [{units[i]}]  ' + '  |  '.join(vals))


if __name__ == '__main__':
    main()
