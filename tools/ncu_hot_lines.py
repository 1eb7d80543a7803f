"""Hottest SASS instructions of a kernel by warp-stall samples, from an .ncu-rep captured with --import-source on.
usage: python tools/ncu_hot_lines.py REPORT [top]"""
import csv, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
hdr = rows[hi[0]]
blk = [r for r in rows[hi[0] + 1:(hi[1] - 1 if len(hi) > 1 else len(rows))] if len(r) >= len(hdr)]
isamp = hdr.index('# Samples')
stalls = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[isamp] or 0) for r in blk)
order = sorted(range(len(blk)), key=lambda i: -int(blk[i][isamp] or 0))[:top]
print('total samples', tot, 'instructions', len(blk))
for i in sorted(order):
    r = blk[i]
    st = sorted(((int(r[j] or 0), hdr[j][6:]) for j in stalls), reverse=True)[:2]
    print(f'{i:5d} {100 * int(r[isamp]) / tot:5.1f}%  {r[1][:70]:70s} ' + ' '.join(f'{n}={v}' for v, n in st if v))
