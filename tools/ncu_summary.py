"""Prints the key counters of an .ncu-rep (raw page) -- used here, on the CPU box, to read profiles."""
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'smsp__cycles_active.avg',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__inst_executed_op_shared_ld.sum', 'smsp__inst_executed_op_shared_st.sum',
        'lts__t_bytes.sum', 'sm__cycles_elapsed.max', 'launch__shared_mem_per_block_dynamic', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'smsp__inst_executed_op_global_ld.sum', 'smsp__inst_executed_op_global_st.sum', 'smsp__inst_executed_op_local_ld.sum',
        'smsp__inst_executed_op_local_st.sum', 'sm__inst_executed_pipe_lsu.sum', 'smsp__inst_issued.sum',
        'sm__sass_inst_executed_op_branch.sum', 'smsp__inst_executed_op_branch.sum']


def main():
    rep = sys.argv[1]
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print(d['Kernel Name'][:100])
        for k in KEYS:
            if k in d:
                print(f'   {k:70s} {d[k]:>20s} {units[hdr.index(k)]}')
        for k in hdr:
            if 'issue_stalled' in k and 'per_issue_active' in k or (len(sys.argv) > 2 and sys.argv[2] in k):
                print(f'   {k:90s} {d[k]}')


if __name__ == '__main__':
    main()
