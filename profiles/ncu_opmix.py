#!/usr/bin/env python
"""Dynamic opcode mix and stall hot spots from an ncu source page:
   ncu -i x.ncu-rep --page source --csv > x_source.csv ; python profiles/ncu_opmix.py x_source.csv [warps]"""
import csv, collections, re, sys
def main(path, nwarps=None, top=40):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    iS, iE, iSamp = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
    ops = collections.Counter(); samp = collections.Counter(); total = 0; tot_s = 0
    lines = []
    for r in rows[2:]:
        if len(r) <= iE: continue
        src = r[iS].strip()
        src = re.sub(r'^@!?U?P\d+\s+', '', src)
        op = src.split()[0].split('.')[0] if src else '?'
        n = int(r[iE] or 0); s = int(r[iSamp] or 0)
        ops[op] += n; samp[op] += s; total += n; tot_s += s
        lines.append((s, n, r[iS].strip()))
    warps = nwarps or max(n for _, n, _ in lines)
    print(f'total warp-instructions {total}; launched warps (max count) {warps}; per warp {total / warps:.0f}')
    for op, n in ops.most_common(top):
        print(f'  {op:14s} {n / warps:9.1f} /warp   samples {100 * samp[op] / max(tot_s, 1):5.1f}%')
if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else None)
