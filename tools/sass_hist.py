#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of a built library: python tools/sass_hist.py lib.so [filter]"""
import collections, re, subprocess, sys
lib = sys.argv[1]; flt = sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, hist = None, {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); hist[cur] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        hist[cur][m.group(1)] += 1
for k, h in hist.items():
    if flt not in k: continue
    tot = sum(h.values())
    print(f"{k}: {tot} instrs")
    print("   " + ", ".join(f"{op}:{c}" for op, c in h.most_common(14)))
