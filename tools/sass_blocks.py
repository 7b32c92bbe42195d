#!/usr/bin/env python
"""Static view of one kernel's SASS (cuobjdump -sass): instruction count, backward branches (loops)
and the opcode mix of any address range.  Used to estimate instructions per cell without a GPU:
    cuobjdump -sass build/k1_fused.o > k1.sass
    python tools/sass_blocks.py k1.sass <mangled-name-substring> [lo:hi[*trips] ...]
With ranges, prints the weighted opcode mix: sum over ranges of (instructions in [lo,hi)) * trips."""
import collections
import re
import sys


def parse(path, key):
    ins = []
    on = False
    for line in open(path):
        if "Function :" in line:
            on = key in line
            continue
        if not on:
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            addr = int(m.group(1), 16)
            txt = m.group(2).strip()
            ins.append((addr, txt))
    return ins


def opcode(txt):
    t = re.sub(r"^@!?U?P\d+\s+", "", txt)
    return t.split()[0].split(".")[0]


def main():
    path, key = sys.argv[1], sys.argv[2]
    ins = parse(path, key)
    print(f"{len(ins)} instructions, {len(ins) * 16 / 1024:.1f} KB")
    for a, t in ins:
        m = re.search(r"\bBRA\b.*?(0x[0-9a-f]+)", t)
        if m and int(m.group(1), 16) <= a:
            print(f"  loop: {int(m.group(1), 16):#x} .. {a:#x}  ({(a - int(m.group(1), 16)) // 16 + 1} instr)  {t}")
        if re.search(r"\b(CALL|RET|EXIT|BSSY|BSYNC)\b", t):
            print(f"  {a:#x}: {t}")
    if len(sys.argv) > 3:
        tot = collections.Counter()
        for spec in sys.argv[3:]:
            rng, _, trips = spec.partition("*")
            lo, hi = (int(x, 16) for x in rng.split(":"))
            k = float(trips) if trips else 1.0
            for a, t in ins:
                if lo <= a < hi:
                    tot[opcode(t)] += k
        n = sum(tot.values())
        fp = tot["DADD"] + tot["DMUL"] + tot["DFMA"]
        print(f"weighted total {n:.0f}; FP64 {fp:.0f}; other {n - fp:.0f}")
        for op, c in tot.most_common(40):
            print(f"  {op:12s} {c:8.0f}")


if __name__ == "__main__":
    main()
