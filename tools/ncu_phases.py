"""Per-phase (between barriers) breakdown of an ncu source-level profile: samples, instructions, stall reasons.
    python tools/ncu_phases.py report.ncu-rep <kernel regex> [launch index]"""
import csv
import io
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat,
                      "--launch-skip", skip, "--launch-count", "1"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
print(rows[0][1][:110])
hdr = rows[1]
ia, isamp, iex = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
ist = {h: i for i, h in enumerate(hdr)}
seen, data = set(), []
for r in rows[2:]:
    if len(r) > 10 and r[isamp].isdigit() and r[0].startswith('0x') and r[0] not in seen:
        seen.add(r[0])
        data.append(r)
print('total samples', sum(int(r[isamp]) for r in data), 'warp instructions', sum(int(r[iex]) for r in data))
keys = (('lsb', 'stall_long_sb'), ('bar', 'stall_barrier'), ('wait', 'stall_wait'), ('ssb', 'stall_short_sb'),
        ('math', 'stall_math'), ('sel', 'stall_selected'), ('lg', 'stall_lg'), ('mio', 'stall_mio'), ('br', 'stall_branch_resolving'))


def new(start):
    d = {'start': start, 'samples': 0, 'inst': 0, 'fp64': 0, 'ld': 0, 'st': 0}
    d.update({k: 0 for k, _ in keys})
    return d


reg, cur = [], new(0)
for pos, r in enumerate(data):
    s = r[ia].strip()
    cur['samples'] += int(r[isamp])
    cur['inst'] += int(r[iex])
    for k, c in keys:
        cur[k] += int(r[ist[c]] or 0)
    op = s.split()[1] if s.startswith('@') else s.split()[0]
    if op.startswith(('DADD', 'DMUL', 'DFMA')):
        cur['fp64'] += int(r[iex])
    if op.startswith(('LDG', 'LDS', 'LDL')):
        cur['ld'] += int(r[iex])
    if op.startswith(('STG', 'STS', 'STL')):
        cur['st'] += int(r[iex])
    if op.startswith(('BAR', 'EXIT')) or 'SYNCS.ARRIVE' in s or 'TRYWAIT' in s:
        cur['end'] = pos
        cur['endop'] = s[:36]
        reg.append(cur)
        cur = new(pos + 1)
reg.append(cur)
for c in reg:
    if c['samples'] > 30 or c['inst'] > 100000:
        print('%5d-%5s samp %6d inst %9d fp64 %8d ld %8d st %7d | ' % (c['start'], c.get('end', 'end'), c['samples'], c['inst'], c['fp64'], c['ld'], c['st'])
              + ' '.join('%s %d' % (k, c[k]) for k, _ in keys if c[k] > 20) + ' | ' + c.get('endop', ''))
