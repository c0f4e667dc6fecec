"""Aggregate an `ncu --page source --csv` dump by opcode: samples, executed warp instructions, shared wavefronts, stalls.
usage: python tools/ncu_src.py dump.csv [kernel-section-index]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
secs = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
start = secs[which]
end = secs[which + 1] if which + 1 < len(secs) else len(rows)
print(rows[start][1][:90])
hdr = rows[start + 1]
idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[start + 2:end] if len(r) == len(hdr)]
tot = sum(int(r[idx['# Samples']] or 0) for r in data)
cls, ex, wf = collections.Counter(), collections.Counter(), collections.Counter()
st = collections.defaultdict(collections.Counter)
keys = ('stall_lg', 'stall_long_sb', 'stall_barrier', 'stall_wait', 'stall_short_sb', 'stall_mio', 'stall_math')
for r in data:
    toks = r[idx['Source']].strip().split()
    op = toks[1] if toks[0].startswith('@') else toks[0]
    op = '.'.join(op.split('.')[:2])
    n = int(r[idx['# Samples']] or 0)
    cls[op] += n
    ex[op] += int(r[idx['Instructions Executed']] or 0)
    wf[op] += int(r[idx['L1 Wavefronts Shared']] or 0)
    for k in keys:
        st[op][k] += int(r[idx[k]] or 0)
print('samples', tot, 'instructions', sum(ex.values()))
for op, n in cls.most_common(int(sys.argv[3]) if len(sys.argv) > 3 else 16):
    print('%-14s %5.1f%% exec %-11d smem_wf %-11d %s' % (op, 100 * n / tot, ex[op], wf[op], {k[6:]: v for k, v in st[op].items() if v}))
