"""Condenses an .ncu-rep (ncu --set full) into the text summary kept under profiles/.

    python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep "title" > profiles/rNN_ncu_x.txt
"""
import csv
import io
import subprocess
import sys

KEEP = ('gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sectors_srcunit_tex_op_read.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.per_cycle_active', 'sm__cycles_elapsed.max', 'dram__throughput.avg.pct_of_peak_sustained_elapsed')


def main():
    rep, title = sys.argv[1], sys.argv[2]
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, units, vals = rows[0], rows[1], rows[2]
    print(f'# ncu --set full --clock-control none, kernel {title}')
    for name in ('Kernel Name', 'Grid Size', 'Block Size'):
        i = head.index(name)
        print(f'{name:106s} {vals[i]}')
    for i, h in enumerate(head):
        if h in KEEP or 'issue_stalled' in h and h.endswith('per_issue_active.ratio'):
            print(f'{h:95s} {units[i]:10s} {vals[i]}')


if __name__ == '__main__':
    main()
