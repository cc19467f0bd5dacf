#!/usr/bin/env python3
"""Counts, per kernel of libbitar_cuda.so, the SASS mnemonics that show what the code uses: bulk-copy engine (UBLKCP /
UBLKPF), mbarrier (SYNCS), warp match / vote / shuffle, shared-memory atomics, vector loads / stores.
usage: python tools/sass_excerpt.py [LIB] > profiles/<tag>_sass_excerpt.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ["UBLKCP", "UBLKPF", "SYNCS", "MATCH", "VOTE", "SHFL", "POPC", "FLO", "ATOMS", "ATOMG", "LDS", "STS", "LDG", "STG", "BAR", "WARPSYNC", "LOP3",
        "SHF", "IMAD", "IADD3", "ISETP", "BRA"]


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "bitar_b200", "csrc", "libbitar_cuda.so")
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    fn, counts, total = None, collections.defaultdict(collections.Counter), collections.Counter()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", line)
        if m and fn:
            op, mods = m.group(1), m.group(2)
            total[fn] += 1
            if op in WANT:
                counts[fn][op] += 1
            if op in ("LDS", "STS", "LDG", "STG") and ".128" in mods:
                counts[fn][op + ".128"] += 1
            if op == "SYNCS" and "TRANS64" in mods:
                counts[fn]["SYNCS.*TRANS64"] += 1
            if op == "MATCH" and ".ANY" in mods:
                counts[fn]["MATCH.ANY"] += 1
    print(f"# SASS mnemonic counts per kernel of {os.path.relpath(lib, ROOT)} (sm_100a cubin, cuobjdump -sass)")
    for fn in sorted(total, key=lambda f: -total[f]):
        print(f"{fn}: {total[fn]} instructions")
        print("   " + "  ".join(f"{k}={v}" for k, v in sorted(counts[fn].items())))


if __name__ == "__main__":
    main()
