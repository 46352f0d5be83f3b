#!/usr/bin/env python
"""Join an ncu SASS source page (ncu -i X.ncu-rep --page source --csv --kernel-name regex:K)
with nvdisasm -g line info to get stall samples per CUDA source line.

usage: sass_line_profile.py <ncu_source.csv> <nvdisasm_-g_listing.txt> <mangled-kernel-substring> [instance]
"""
import collections
import csv
import re
import sys


def main():
    src_csv, listing, kern = sys.argv[1:4]
    inst = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    # --- nvdisasm: ordered (line) per instruction of the kernel's .text section
    lines = open(listing, errors="replace").read().split("\n")
    start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and kern in l and l.rstrip().endswith(":"))
    cur = None
    per_instr = []
    for l in lines[start + 1:]:
        if l.startswith("//-----") or l.startswith(".section") and per_instr:
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4}\*/", l):
            per_instr.append(cur)
    # --- ncu: rows of the requested kernel instance
    rows = list(csv.reader(open(src_csv)))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    s = starts[inst]
    e = starts[inst + 1] if inst + 1 < len(starts) else len(rows)
    hdr = rows[s + 1]
    body = rows[s + 2:e]
    ci = hdr.index("# Samples")
    ii = hdr.index("Instructions Executed")
    print(f"instructions: ncu={len(body)} nvdisasm={len(per_instr)}")
    agg = collections.Counter()
    aggi = collections.Counter()
    n = min(len(body), len(per_instr))
    for k in range(n):
        agg[per_instr[k]] += int(body[k][ci] or 0)
        aggi[per_instr[k]] += int(body[k][ii] or 0)
    tot = sum(agg.values())
    print(f"total samples {tot}")
    for key, v in agg.most_common(45):
        print(f"{v:7d} {100.0 * v / tot:5.1f}%  inst={aggi[key]:8d}  {key}")


if __name__ == "__main__":
    main()
