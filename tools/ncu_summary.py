#!/usr/bin/env python
"""Text summary of an .ncu-rep (ncu --set full): the launch, pipe and memory metrics the DESIGN / profiles notes quote,
plus stall reasons per issued instruction.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [title] > profiles/rN_ncu_summary.txt
"""
import csv
import subprocess
import sys

METRICS = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
    'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
    'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
    'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'sm__cycles_elapsed.max', 'sm__cycles_active.avg',
    'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
]
STALL = 'smsp__average_warps_issue_stalled_'


def main():
    rep = sys.argv[1]
    if len(sys.argv) > 2:
        print(sys.argv[2])
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print(r[ix['Kernel Name']][:150])
        for m in METRICS:
            if m in ix and r[ix[m]] != '':
                print(f'   {m} {r[ix[m]]} {units[ix[m]]}')
        stalls = []
        for h in hdr:
            if h.startswith(STALL) and h.endswith('_per_warp_active.pct') is False and h.endswith('.ratio'):
                try:
                    stalls.append((float(r[ix[h]]), h[len(STALL):].replace('_per_issue_active.ratio', '').replace('.ratio', '')))
                except ValueError:
                    pass
        if stalls:
            stalls.sort(reverse=True)
            print('   stalls per issue: ' + ' '.join(f'{n}={v:.2f}' for v, n in stalls[:10]))


if __name__ == '__main__':
    main()
