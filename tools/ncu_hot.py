#!/usr/bin/env python
"""Hot spots from an `ncu --page source --csv` dump: stall totals, top instructions, per-opcode samples."""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1]))]
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
data = []
for r in rows[hi + 1:]:
    if not r or r[0] in ("Address", "Kernel Name"):
        break                      # first captured launch only
    if len(r) == len(hdr):
        data.append(r)
idx = {h: i for i, h in enumerate(hdr)}
S, E = idx['# Samples'], idx['Instructions Executed']
tot = sum(int(r[S]) for r in data)
print('total samples', tot, 'sass instrs', len(data), 'warp-instr executed', sum(int(r[E]) for r in data))
st = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {h: sum(int(r[idx[h]] or 0) for r in data) for h in st}
print('stalls:', ', '.join('%s=%.1f%%' % (k[6:], 100.0 * v / max(tot, 1)) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:9]))
for r in sorted(data, key=lambda r: -int(r[S]))[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    s = {h: int(r[idx[h]] or 0) for h in st}
    dom = max(s.items(), key=lambda kv: kv[1])
    print(r[S].rjust(6), r[E].rjust(9), dom[0][6:].ljust(14), r[idx['Source']].strip()[:90])
c, ce = collections.Counter(), collections.Counter()
for r in data:
    m = re.match(r'\s*(@!?U?P\w+\s+)?([A-Z0-9_]+)', r[idx['Source']])
    op = m.group(2) if m else '?'
    c[op] += int(r[S])
    ce[op] += int(r[E])
print('by opcode (samples, executed):', [(k, v, ce[k]) for k, v in c.most_common(14)])
