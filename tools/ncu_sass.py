#!/usr/bin/env python
"""SASS-level digest of one kernel in an .ncu-rep: opcode mix per thread and the instructions with the
most stall samples.  usage: ncu_sass.py REP KERNEL_REGEX [units_for_normalisation] [--top N] [--dump]"""
import collections
import csv
import io
import re
import subprocess
import sys


def main():
    rep, rx = sys.argv[1], sys.argv[2]
    units = float(sys.argv[3]) if len(sys.argv) > 3 and not sys.argv[3].startswith("--") else 1.0
    top_n = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 25
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    # split per kernel section
    sections, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            sections.append(cur)
        elif cur is not None:
            cur["rows"].append(r)
    for sec in sections:
        if not re.search(rx, sec["name"]):
            continue
        hdr, data = sec["rows"][0], sec["rows"][1:]
        ix = {h: i for i, h in enumerate(hdr)}
        tot = sum(int(r[ix["Instructions Executed"]]) for r in data)
        samples = sum(int(r[ix["# Samples"]]) for r in data)
        print(f"== {sec['name'][:100]}")
        print(f"   SASS instructions {len(data)}, warp-instructions executed {tot}, per unit {tot * 32 / units:.2f} thread-instr, samples {samples}")
        ops, st = collections.Counter(), collections.Counter()
        for r in data:
            src = r[ix["Source"]].split()
            op = (src[1] if src[0].startswith("@") else src[0]).split(".")[0]
            ops[op] += int(r[ix["Instructions Executed"]])
            st[op] += int(r[ix["# Samples"]])
        print("   opcode mix (thread-instr per unit | share of stall samples):")
        for op, n in ops.most_common(18):
            print(f"     {op:10s} {n * 32 / units:8.2f}   {100.0 * st[op] / max(samples, 1):5.1f}%")
        print("   hottest instructions by stall samples:")
        order = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]]))[:top_n]
        for i in sorted(order):
            r = data[i]
            print(f"     #{i:5d} {100.0 * int(r[ix['# Samples']]) / max(samples, 1):5.1f}%  exec {int(r[ix['Instructions Executed']]):>10d}  {r[ix['Source']].strip()[:80]}")
        if "--dump" in sys.argv:
            for i, r in enumerate(data):
                print(f"{i:5d} {int(r[ix['Instructions Executed']]):>10d} {int(r[ix['# Samples']]):>6d}  {r[ix['Source']].strip()}")
        break


if __name__ == "__main__":
    main()
