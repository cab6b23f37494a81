"""SASS fingerprint of libabismal_b200.so: per kernel the resource usage (cuobjdump -res-usage) and the
histogram of instruction mnemonics (cuobjdump -sass), plus one hash over the instruction stream.  Committed
under profiles/ every round so that "machine code unchanged" claims can be checked.
usage: sass_fingerprint.py [library] > profiles/rNN_sass_fingerprint.txt"""
import collections
import hashlib
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "abismal_b200", "libabismal_b200.so")
res = subprocess.run(["cuobjdump", "-res-usage", lib], stdout=subprocess.PIPE, text=True).stdout
sass = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True).stdout
print("# %s" % os.path.relpath(lib, ROOT))
usage = {}
name = None
for ln in res.splitlines():
    m = re.search(r"Function (\S+):", ln)
    if m:
        name = m.group(1)
    elif name and "REG:" in ln:
        usage[name] = ln.strip()
        name = None
cur, hist, stream = None, {}, {}
for ln in sass.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        hist[cur] = collections.Counter()
        stream[cur] = hashlib.sha256()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", ln)
    if cur and m:
        ins = m.group(1)
        ins = re.sub(r"^@!?U?P\d+\s+", "", ins)
        op = ins.split()[0]
        hist[cur][op] += 1
        stream[cur].update(m.group(1).encode())
for fn in sorted(hist):
    dem = subprocess.run(["cu++filt", fn], stdout=subprocess.PIPE, text=True).stdout.strip() or fn
    n = sum(hist[fn].values())
    print("\n== %s\n   %s\n   %s\n   instructions %d (%.1f KB), sha256 %s" % (dem, fn, usage.get(fn, "?"), n, n * 16 / 1024.0,
                                                                              stream[fn].hexdigest()[:16]))
    groups = collections.Counter()
    for op, c in hist[fn].items():
        groups[op.split(".")[0]] += c
    print("   " + "  ".join("%s %d" % kv for kv in groups.most_common(24)))
    for key in ("LDG", "STG", "LDS", "STS", "LDL", "STL", "SHFL", "POPC", "ATOM", "RED", "BAR", "VOTE", "LDGSTS"):
        det = sorted((op, c) for op, c in hist[fn].items() if op.split(".")[0] == key)
        if det:
            print("   %s: %s" % (key, ", ".join("%s x%d" % kv for kv in det)))
