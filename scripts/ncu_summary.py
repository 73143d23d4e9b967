#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here on the CPU box with `ncu -i`) into a small
text table for profiles/: per-launch duration, DRAM traffic, pipe/issue
utilisation, occupancy, shared-memory wavefronts and bank conflicts."""
import csv
import subprocess
import sys

METRICS = [
    ('gpu__time_duration.sum', 'time'),
    ('dram__bytes_read.sum', 'dram_rd'),
    ('dram__bytes_write.sum', 'dram_wr'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram%'),
    ('launch__registers_per_thread', 'regs'),
    ('launch__grid_size', 'grid'),
    ('launch__block_size', 'block'),
    ('launch__shared_mem_per_block_dynamic', 'dyn_smem'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps_active%'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue_active%'),
    ('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'fma_pipe%'),
    ('sm__inst_executed_pipe_lsu.sum', 'lsu_inst'),
    ('smsp__inst_executed.sum', 'warp_inst'),
    ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smem_wavefronts'),
    ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smem_bank_conflicts'),
    ('l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex%'),
    ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l2%'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm%'),
]


def main(path):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print('# %s' % path)
    for r in rows[2:]:
        print('\n## %s  (id %s)' % (r[idx['Kernel Name']], r[idx['ID']]))
        for m, short in METRICS:
            if m in idx:
                print('%-22s %s %s' % (short, r[idx[m]], units[idx[m]]))


if __name__ == '__main__':
    main(sys.argv[1])
