#!/usr/bin/env python
"""Summarise an .ncu-rep (read with `ncu -i` here on the CPU box): headline metrics + stall-reason attribution.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--top 6] [--json out.json]
"""
import csv
import io
import json
import subprocess
import sys

METRICS = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__warps_active.avg.per_cycle_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
    'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_lsu.sum',
    'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
    'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum',
    'sm__cycles_elapsed.max', 'smsp__cycles_active.avg',
]


def ncu_csv(rep, page):
    out = subprocess.run(['ncu', '-i', rep, '--page', page, '--csv'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                         text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index('--top') + 1]) if '--top' in sys.argv else 5
    raw = ncu_csv(rep, 'raw')
    hdr, units, rows = raw[0], raw[1], raw[2:]
    summary = []
    for r in rows:
        rec = {'kernel': r[hdr.index('Kernel Name')]}
        for m in METRICS:
            if m in hdr:
                rec[m] = (r[hdr.index(m)], units[hdr.index(m)])
        stalls = {}
        for i, h in enumerate(hdr):
            if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio'):
                try:
                    stalls[h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]] = float(r[i])
                except ValueError:
                    pass
        rec['stalls_per_issue'] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:10])
        summary.append(rec)
    for rec in summary:
        print('=== ', rec['kernel'][:110])
        for m in METRICS:
            if m in rec:
                print(f'  {m:72s} {rec[m][0]:>16s} {rec[m][1]}')
        print('  stalls per issue:', ', '.join(f'{k}={v:.2f}' for k, v in rec['stalls_per_issue'].items()))
    src = ncu_csv(rep, 'source')
    hi = next(i for i, r in enumerate(src) if r and r[0] == 'Address')
    shdr = src[hi]
    ix = {h: i for i, h in enumerate(shdr)}
    first = []
    for r in src[hi + 1:]:
        if r and r[0] == 'Kernel Name':
            break
        if len(r) == len(shdr):
            first.append(r)
    tot = sum(int(r[ix['# Samples']]) for r in first)
    print(f'--- source page, first kernel: {len(first)} SASS instructions, {tot} stall samples')
    keys = [k for k in shdr if k.startswith('stall_') and '(Not Issued)' not in k]
    agg = sorted(((sum(int(r[ix[k]]) for r in first), k) for k in keys), reverse=True)
    for s, k in agg[:9]:
        print(f'  {k:26s} {s:7d} ({100.0 * s / max(tot, 1):5.1f}%)')
        for r in sorted(first, key=lambda r: -int(r[ix[k]]))[:top]:
            if int(r[ix[k]]) > 0:
                print(f'        {r[ix[k]]:>6s}  {r[ix["Source"]][:96]}')
    if '--json' in sys.argv:
        with open(sys.argv[sys.argv.index('--json') + 1], 'w') as f:
            json.dump(summary, f, indent=1)


if __name__ == '__main__':
    main()
