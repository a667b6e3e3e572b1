#!/usr/bin/env python
"""Hot SASS instructions of one kernel in an .ncu-rep (needs --import-source on / -lineinfo).
    python tools/ncu_hot.py <rep> <kernel-regex> [topN]"""
import csv, io, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', f'regex:{rx}'], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith('"Kernel Name"')), len(lines))
rows = list(csv.DictReader(io.StringIO('\n'.join(lines[start:end]))))
stall_cols = [c for c in rows[0].keys() if c.startswith('stall_') and 'Not Issued' not in c]
tot = sum(int(r['# Samples'] or 0) for r in rows)
print(f'{len(rows)} SASS instructions, {tot} samples')
agg = {c: sum(int(r[c] or 0) for r in rows) for c in stall_cols}
print('stall totals:', ', '.join(f'{k[6:]} {100 * v / max(tot, 1):.1f}%' for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
for i, r in enumerate(rows):
    r['_i'] = i
for r in sorted(rows, key=lambda r: -int(r['# Samples'] or 0))[:top]:
    n = int(r['# Samples'] or 0)
    st = sorted(((c[6:], int(r[c] or 0)) for c in stall_cols), key=lambda kv: -kv[1])[:2]
    print(f"{r['_i']:5d} {100 * n / tot:5.1f}%  exec {r['Instructions Executed']:>9}  {st[0][0]}:{st[0][1]} {st[1][0]}:{st[1][1]}  {r['Source'][:110]}")
