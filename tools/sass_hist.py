#!/usr/bin/env python
"""Opcode histogram of a kernel's hot loop, from `cuobjdump -sass` (no GPU needed).

    python tools/sass_hist.py <lib.so | file.cubin | file.o> <kernel-name-regex> [rows-per-iteration]

The hot loop is taken to be the longest backward branch without an atomic or an exit inside (for the score
kernels: the five-row unrolled body, not the work-queue loop around it).  Prints instructions per iteration and per row by opcode, spill instructions
inside the loop, and the share of FADD / FMNMX(3) -- the numbers DESIGN.md quotes per DP row.
"""
import collections
import re
import subprocess
import sys


def functions(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    name, body = None, []
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if name:
                yield name, body
            name, body = m.group(1), []
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and name:
            body.append((int(m.group(1), 16), m.group(2).strip()))
    if name:
        yield name, body


def main():
    path, pat = sys.argv[1], re.compile(sys.argv[2])
    rows = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    for name, body in functions(path):
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        if not pat.search(dem) and not pat.search(name):
            continue
        best = None
        for addr, ins in body:
            m = re.search(r"\bBRA\s+(?:`\(\S+\)|0x([0-9a-f]+))", ins)
            if m and m.group(1):
                tgt = int(m.group(1), 16)
                inner = [i for a, i in body if tgt <= a <= addr]
                # the row loop, not the work-queue loop around it: no atomics, no exit inside
                if any(re.search(r"\b(ATOMG|ATOM|EXIT|RED)\b", i) for i in inner):
                    continue
                if tgt < addr and (best is None or addr - tgt > best[1] - best[0]):
                    best = (tgt, addr)
        print("== %s" % dem[:150])
        if not best:
            print("   no backward branch found")
            continue
        loop = [ins for addr, ins in body if best[0] <= addr <= best[1]]
        hist = collections.Counter()
        for ins in loop:
            t = ins.split()
            op = t[1] if t[0].startswith("@") else t[0]
            hist[op.split(".")[0] + ("." + op.split(".")[1] if op.startswith(("LDG", "STG", "LDS", "STS", "SHFL")) and "." in op else "")] += 1
        n = len(loop)
        print("   loop 0x%x..0x%x: %d instructions per iteration = %.1f per row (%d rows per iteration); function %d instructions"
              % (best[0], best[1], n, n / rows, rows, len(body)))
        spills = sum(v for k, v in hist.items() if k.startswith(("STL", "LDL")))
        print("   spill instructions inside the loop: %d" % spills)
        for op, c in hist.most_common():
            print("   %-14s %5d  %6.1f / row  %5.1f %%" % (op, c, c / rows, 100.0 * c / n))


if __name__ == "__main__":
    main()
