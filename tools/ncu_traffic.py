#!/usr/bin/env python3
"""Summarises an `ncu --set full` report of tools/gpu_roundtrip_once.py into the two files the round commits:
   profiles/<tag>_ncu_summary.txt  key counters per kernel (time, instructions, issue utilisation, stalls, bank conflicts, DRAM bytes)
   profiles/r02_dram_traffic.json  DRAM bytes per uncompressed byte of the deflate / inflate kernel (bench.py's roofline.traffic)
usage: python tools/ncu_traffic.py REPORT.ncu-rep TAG UNCOMPRESSED_BYTES"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'smsp__cycles_active.avg', 'sm__cycles_elapsed.max',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct']
SCALE = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0}


def main():
    rep, tag, nbytes = sys.argv[1], sys.argv[2], float(sys.argv[3])
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    lines, traffic = [], {}
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        name = d['Kernel Name']
        lines.append(name[:110])
        for k in KEYS:
            if k in d:
                lines.append(f'   {k:78s} {d[k]:>20s} {units[hdr.index(k)]}')
        for k in hdr:
            if 'issue_stalled' in k and 'per_issue_active' in k:
                lines.append(f'   {k:98s} {d[k]}')
        dram = sum(float(d[k]) * SCALE.get(units[hdr.index(k)], 1.0) for k in ('dram__bytes_read.sum', 'dram__bytes_write.sum'))
        lines.append(f'   DRAM bytes (read + write) per uncompressed byte of the launch: {dram / nbytes:.3f}')
        key = 'deflate_bytes_per_byte' if 'deflate' in name else 'inflate_bytes_per_byte' if 'inflate_tok' in name else None
        if key:
            traffic[key] = dram / nbytes
    with open(os.path.join(ROOT, 'profiles', f'{tag}_ncu_summary.txt'), 'w') as f:
        f.write('\n'.join(lines) + '\n')
    if len(traffic) == 2 and 'seg59460' in tag:   # bench.py's roofline.traffic is quoted on the benchmark's segment size
        traffic['source'] = (f'profiles/{tag}_ncu_summary.txt: dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of '
                             f'tools/gpu_roundtrip_once.py over {int(nbytes)} bytes, per uncompressed byte')
        with open(os.path.join(ROOT, 'profiles', 'r02_dram_traffic.json'), 'w') as f:
            json.dump(traffic, f, indent=1)
    print('\n'.join(lines))


if __name__ == '__main__':
    main()
