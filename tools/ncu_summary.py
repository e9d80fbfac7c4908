#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, without a GPU): key raw metrics per kernel + hottest source lines.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [n_lines]"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_sector_hit_rate.pct', 'sm__cycles_elapsed.avg', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed']


def run(args):
    return subprocess.run(['ncu'] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    nlines = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    rows = list(csv.reader(io.StringIO(run(['-i', rep, '--page', 'raw', '--csv']))))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print('=== kernel', d.get('Kernel Name', '?')[:60], 'id', d.get('ID'))
        for k in KEYS:
            if k in d:
                print(f'  {k:75s} {d[k]:>16s} {units[hdr.index(k)]}')
        st = [(h.replace('smsp__pcsamp_warps_issue_stalled_', ''), int(float(d[h] or 0))) for h in hdr
              if 'pcsamp_warps_issue_stalled' in h and 'not_issued' not in h]
        tot = sum(v for _, v in st) or 1
        print('  stalls:', ', '.join(f'{k} {100 * v / tot:.0f}%' for k, v in sorted(st, key=lambda kv: -kv[1])[:8]))
    src = list(csv.reader(io.StringIO(run(['-i', rep, '--page', 'source', '--csv', '--print-source', 'sass,cuda']))))
    cur = func = None
    h = None
    agg = collections.defaultdict(lambda: [0, 0, ''])
    for r in src:
        if len(r) == 2 and r[0] == 'File Path':
            cur = r[1]
            continue
        if len(r) == 2 and r[0] == 'Function Name':
            func = r[1][:40]
            continue
        if r and r[0] == 'Line No':
            h = r
            continue
        if h and len(r) == len(h):
            d = dict(zip(h, r))
            try:
                ln = int(r[0])
            except ValueError:
                continue
            key = (func, (cur or '?').split('/')[-1], ln)
            agg[key][0] += int(float(d.get('Instructions Executed', '0') or 0))
            agg[key][1] += int(float(d.get('# Samples', '0') or 0))
            agg[key][2] = r[1][:100].strip()
    tot = collections.Counter()
    tots = collections.Counter()
    for k, v in agg.items():
        tot[k[0]] += v[0]
        tots[k[0]] += v[1]
    for f in tot:
        print('== source', f, tot[f], 'warp instr', tots[f], 'samples')
        items = sorted(((k, v) for k, v in agg.items() if k[0] == f), key=lambda kv: -kv[1][0])
        for k, v in items[:nlines]:
            print(f'{k[1]}:{k[2]:4d} inst {100 * v[0] / max(1, tot[f]):5.1f}% samp {100 * v[1] / max(1, tots[f]):5.1f}%  {v[2]}')


if __name__ == '__main__':
    main()
