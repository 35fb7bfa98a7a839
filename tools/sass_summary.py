"""Opcode summary of every kernel in the built library (cuobjdump -sass): what the hardware is asked to do.

  python tools/sass_summary.py dataassimbench_b200/lib/libdab_b200.so profiles/sass_r02.md
"""
import collections
import re
import subprocess
import sys

KEY = ['DMMA', 'DFMA', 'DADD', 'DMUL', 'LDG', 'STG', 'LDS', 'STS', 'LDGSTS', 'UBLKCP', 'UTMALDG', 'UTMASTG', 'SYNCS', 'BAR',
       'UCGABAR', 'ACQBULK', 'PREEXIT', 'ATOM', 'ATOMG', 'RED', 'SHFL', 'MUFU', 'MEMBAR', 'ERRBAR', 'CCTL']


def main(lib, dst):
    out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True, check=True).stdout
    rows = []
    for block in re.split(r'\n\s*Function : ', out)[1:]:
        name = block.split('\n', 1)[0].strip()
        dem = subprocess.run(['cu++filt', name], capture_output=True, text=True).stdout.strip() or name
        dem = re.sub(r'^void ', '', dem).replace('(int)', '').replace('(bool)', '')
        dem = re.sub(r'\(.*', '', dem).replace('dab::', '').replace('(anonymous namespace)::', '')
        ops = collections.Counter()
        uniform = 0
        n = 0
        for line in block.split('\n'):
            m = re.match(r'\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)\s*(.*?);', line)
            if not m:
                continue
            n += 1
            ops[m.group(1)] += 1
            if m.group(1) in ('DFMA', 'DADD', 'DMUL') and re.search(r'\bUR\d+', m.group(3)):
                uniform += 1
        rows.append((dem, n, ops, uniform))
    rows.sort(key=lambda r: r[0])
    with open(dst, 'w') as f:
        f.write('# SASS opcode summary of `%s` (sm_100a)\n\n' % lib)
        f.write('`cuobjdump -sass`, static instruction counts per kernel (`tools/sass_summary.py`).  `DMMA` = `DMMA.8x8x4`, the FP64 '
                'tensor-core instruction (`tcgen05` has no FP64 kind); `UBLKCP` = `cp.async.bulk` (multicast in K4); `LDGSTS` = '
                '`cp.async`; `ACQBULK`/`PREEXIT` = programmatic dependent launch; `UCGABAR` = cluster barrier; "uniform" = FP64 '
                'instructions with a uniform-register operand (full issue rate, see profiles/forecast_probes_r02.md section 1).\n\n')
        f.write('| kernel | instr | ' + ' | '.join(KEY) + ' | FP64 w/ uniform operand |\n')
        f.write('|---|---|' + '---|' * len(KEY) + '---|\n')
        for dem, n, ops, uniform in rows:
            f.write('| `%s` | %d | ' % (dem[:90], n) + ' | '.join(str(ops.get(k_, 0) or '') for k_ in KEY) + ' | %s |\n' % (uniform or ''))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2])
