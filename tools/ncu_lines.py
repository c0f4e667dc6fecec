"""Aggregate the pc samples of an `ncu --page source --csv` dump by CUDA source line (line table from `nvdisasm -g -c`
of the same cubin).  usage: python tools/ncu_lines.py dump.csv nvdisasm.txt <mangled-kernel-name> [top [section]]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
secs = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
which = int(sys.argv[5]) if len(sys.argv) > 5 else 0
start = secs[which]
end = secs[which + 1] if which + 1 < len(secs) else len(rows)
hdr = rows[start + 1]
idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[start + 2:end] if len(r) == len(hdr)]
lines, cur, infn = [], None, False
for line in open(sys.argv[2]):
    if '.section' in line and '.text.' in line:
        infn = sys.argv[3] in line
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', line)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)), (m.group(3) or '').split('/')[-1], int(m.group(4) or 0))
        continue
    if re.match(r'\s+/\*[0-9a-f]+\*/', line):
        lines.append(cur)
assert len(lines) == len(data), (len(lines), len(data))
keys = [k for k in hdr if k.startswith('stall_')]
tot = sum(int(r[idx['# Samples']] or 0) for r in data)
agg = collections.Counter()
st = collections.defaultdict(collections.Counter)
for ln, r in zip(lines, data):
    n = int(r[idx['# Samples']] or 0)
    key = (ln[2], ln[3]) if ln and ln[2] else (ln[0], ln[1]) if ln else ('?', 0)
    agg[key] += n
    for k in keys:
        st[key][k[6:]] += int(r[idx[k]] or 0)
print('samples', tot)
for key, n in agg.most_common(int(sys.argv[4]) if len(sys.argv) > 4 else 30):
    print('%-22s %5d %5.1f%% %s' % (key[0], key[1], 100 * n / tot, dict(st[key].most_common(3))))
