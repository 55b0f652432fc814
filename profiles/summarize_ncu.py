#!/usr/bin/env python3
'''Turns an `ncu --set full` report (gpurun_out/*.ncu-rep, scratch) into the small CSV summaries kept in
profiles/:   python profiles/summarize_ncu.py REPORT.ncu-rep "comment line" "command line" > profiles/NAME.csv

With --traffic KEY NAME.csv as trailing arguments the DRAM bytes of the launch (read + write) are also recorded in
profiles/traffic.json under KEY ('<problem>_<grid>_<N>gpu'), which is where bench.py takes `roofline.traffic` from.'''
import json
import os
import csv
import subprocess
import sys

KEEP = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
    'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
    'launch__occupancy_limit_shared_mem', 'launch__shared_mem_per_block_dynamic',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
    'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
]


def main():
    rep, comment, command = sys.argv[1], sys.argv[2], sys.argv[3]
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, val = rows[0], rows[1], rows[2]
    print('# ' + comment)
    print('# command: ' + command)
    if '--all' in sys.argv:
        # one block per captured launch (a report that holds several kernels)
        ki = hdr.index('Kernel Name')
        for val in rows[2:]:
            print('## kernel: ' + val[ki])
            stalls = []
            for h, u, v in zip(hdr, units, val):
                if h in KEEP:
                    print('%s,%s,%s' % (h, u, v))
                elif 'issue_stalled' in h and h.endswith('per_issue_active.ratio'):
                    stalls.append((float(v.replace(',', '')), h))
            for v, h in sorted(stalls, reverse=True)[:5]:
                print('%s,ratio,%.3f' % (h, v))
        return
    stalls = []
    for h, u, v in zip(hdr, units, val):
        if h in KEEP:
            print('%s,%s,%s' % (h, u, v))
        elif 'issue_stalled' in h and h.endswith('per_issue_active.ratio'):
            stalls.append((float(v.replace(',', '')), h))
    for v, h in sorted(stalls, reverse=True)[:8]:
        print('%s,ratio,%.3f' % (h, v))
    if '--traffic' in sys.argv:
        key, source = sys.argv[sys.argv.index('--traffic') + 1], sys.argv[sys.argv.index('--traffic') + 2]
        scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
        byts = {}
        for h, u, v in zip(hdr, units, val):
            if h in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
                byts[h] = float(v.replace(',', '')) * scale[u]
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'traffic.json')
        table = {}
        if os.path.exists(path):
            with open(path) as f:
                table = json.load(f)
        table[key] = {'dram_bytes_per_launch': sum(byts.values()), 'dram_bytes_read': byts.get('dram__bytes_read.sum'),
                      'dram_bytes_write': byts.get('dram__bytes_write.sum'), 'source': 'profiles/' + os.path.basename(source),
                      'comment': comment}
        with open(path, 'w') as f:
            json.dump(table, f, indent=1, sort_keys=True)


if __name__ == '__main__':
    main()
