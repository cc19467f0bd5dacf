#!/usr/bin/env python3
"""Per-source-line summary of an `ncu --page source --csv --print-source cuda,sass` dump.

usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:NAME > dump.csv
       python tools/ncu_lines.py dump.csv [top_n] [file-substring]
Prints, for the first kernel instance of the dump, the lines with the most executed warp instructions and the
most stall samples (instructions, share, samples, share, dominant stall reasons).
"""
import csv
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    filt = sys.argv[3] if len(sys.argv) > 3 else ""
    cur_file, hdr, line = None, None, None
    seen_func = set()
    inst = defaultdict(int)
    samp = defaultdict(int)
    reasons = defaultdict(lambda: defaultdict(int))
    text = {}
    first_func = None
    skip = False
    for r in csv.reader(open(path)):
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1]
            continue
        if r[0] == "Function Name":
            key = (cur_file, r[1])
            if first_func is None:
                first_func = r[1]
            skip = key in seen_func or r[1] != first_func
            seen_func.add(key)
            continue
        if r[0] == "Line No":
            hdr = {n: i for i, n in enumerate(r)}
            ncol = len(r)
            ri = [(n, i) for i, n in enumerate(r) if n.startswith("stall_") and "Not Issued" not in n]
            continue
        if skip or hdr is None:
            continue
        if r[0] == "":
            continue          # the SASS rows under a line (listed twice); the line's own row carries the sums
        line = (cur_file.split("/")[-1], int(r[0]))
        text[line] = r[1].strip()
        sh = len(r) - ncol      # ncu does not escape quotes inside the source text: such rows have extra columns
        try:
            inst[line] += int(r[hdr["Instructions Executed"] + sh])
            samp[line] += int(r[hdr["# Samples"] + sh])
            for n, i in ri:
                v = int(r[i + sh])
                if v:
                    reasons[line][n[6:]] += v
        except (ValueError, IndexError):
            pass
    # inlined callees appear once per level of the inline stack: totals over the kernel's own files only
    own = [k for k in inst if not k[0].endswith(".hpp")]
    ti, ts = sum(inst[k] for k in own) or 1, sum(samp[k] for k in own) or 1
    print(f"kernel: {first_func[:80]}\nwarp instructions {ti}  samples {ts}")
    keys = [k for k in inst if filt in k[0]]
    print("\n== by executed warp instructions ==")
    for k in sorted(keys, key=lambda k: -inst[k])[:top]:
        print(f"{k[0]}:{k[1]:<5d} inst {inst[k]:>11d} {100*inst[k]/ti:5.1f}%  samp {100*samp[k]/ts:5.1f}%  {text.get(k,'')[:100]}")
    print("\n== by stall samples ==")
    for k in sorted(keys, key=lambda k: -samp[k])[:top]:
        rs = sorted(reasons[k].items(), key=lambda x: -x[1])[:3]
        print(f"{k[0]}:{k[1]:<5d} samp {100*samp[k]/ts:5.1f}%  inst {100*inst[k]/ti:5.1f}%  {' '.join(f'{n}={v}' for n,v in rs):40s} {text.get(k,'')[:80]}")


if __name__ == "__main__":
    main()
