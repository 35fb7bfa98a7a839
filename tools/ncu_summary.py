"""Turn gpurun_out/*.ncu-rep / launch lists into the small text summaries committed under profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches_r01.csv profiles/launches_r01_summary.md
  python tools/ncu_summary.py kernel gpurun_out/prof_forecast_r01.ncu-rep profiles/forecast_r01_ncu.md
  python tools/ncu_summary.py kernel gpurun_out/prof_analysis_r01.ncu-rep profiles/ns_r01_ncu.md ns_transform
"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__cycles_active.avg', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'dram__bytes_read.sum',
        'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    idx = {h: i for i, h in enumerate(rows[0])}
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name = re.sub(r'^void ', '', re.sub(r'\(.*', '', r[idx['Kernel Name']]))
        val = float(r[idx['Metric Value']].replace(',', ''))
        unit = r[idx['Metric Unit']]
        val = val / 1e3 if unit.startswith('n') else (val * 1e3 if unit.startswith('m') else val)
        a = agg.setdefault((name, r[idx['Block Size']], r[idx['Grid Size']]), [0, 0.0])
        a[0] += 1
        a[1] += val
    tot = sum(v[1] for v in agg.values())
    with open(dst, 'w') as f:
        f.write('# ncu launch list summary (%s)\n\n' % os.path.basename(src))
        f.write('`ncu --metrics gpu__time_duration.sum --clock-control none` over a short `python bench.py '
                '--no-cpu-baseline` run (see tools/evidence_r01.sh); per-launch times are cold-cache and serialised: compare SHARES.\n\n')
        f.write('| kernel | block | grid | launches | avg us | share |\n|---|---|---|---|---|---|\n')
        for (name, blk, grd), v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write('| `%s` | %s | %s | %d | %.1f | %.3f |\n' % (name[:80], blk, grd, v[0], v[1] / v[0], v[1] / tot))


def kernel(src, dst, name_filter=None):
    flt = ['-k', 'regex:' + name_filter] if name_filter else []
    raw = subprocess.run(['ncu', '-i', src] + flt + ['--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    got = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
    srcp = subprocess.run(['ncu', '-i', src] + flt + ['--page', 'source', '--csv'], capture_output=True, text=True).stdout
    srows = list(csv.reader(srcp.splitlines()))
    name = srows[0][1]
    sh = srows[1]
    si = {h: i for i, h in enumerate(sh)}
    data = []
    for r in srows[2:]:                     # first kernel section only
        if len(r) < len(sh) or r[0] == 'Kernel Name':
            break
        data.append(r)
    tot = sum(int(r[si['# Samples']]) for r in data) or 1
    stalls = [h for h in sh if h.startswith('stall_') and 'Not Issued' not in h]
    agg = sorted(((sum(int(r[si[s]]) for r in data) / tot, s) for s in stalls), reverse=True)
    ops = collections.Counter()
    for r in data:
        op = re.sub(r'^@!?U?P\d+\s+', '', r[si['Source']].strip()).split(' ')[0].split('.')[0]
        ops[op] += int(r[si['Instructions Executed']] or 0)
    with open(dst, 'w') as f:
        f.write('# ncu --set full summary: `%s`\n\nsource: `%s` (scratch, not committed)\n\n' % (name, src))
        f.write('| metric | value | unit |\n|---|---|---|\n')
        for k in KEYS:
            if k in got:
                f.write('| %s | %s | %s |\n' % (k, got[k][0], got[k][1]))
        f.write('\nWarp stall sampling (%d samples), share of all samples:\n\n' % tot)
        for v, s in agg:
            if v >= 0.005:
                f.write('* %s: %.1f %%\n' % (s, 100 * v))
        f.write('\nExecuted warp instructions by opcode (top 12): ' +
                ', '.join('%s %d' % kv for kv in ops.most_common(12)) + '\n')
    rd = float(got['dram__bytes_read.sum'][0].replace(',', ''))
    wr = float(got['dram__bytes_write.sum'][0].replace(',', ''))
    mult = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    return rd * mult[got['dram__bytes_read.sum'][1]] + wr * mult[got['dram__bytes_write.sum'][1]]


if __name__ == '__main__':
    if sys.argv[1] == 'launches':
        launches(sys.argv[2], sys.argv[3])
    else:
        t = kernel(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
        print(json.dumps({'dram_bytes_per_launch': t}))
