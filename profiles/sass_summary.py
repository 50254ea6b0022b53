"""Per-kernel opcode histogram of lib/libswrt.so (sm_100a SASS) -- the evidence for which hardware mechanisms the binary uses:
UTMALDG (TMA bulk-tensor loads), SYNCS (mbarrier), LDGSTS (cp.async), SHFL (warp shuffles), DFMA/DADD/DMUL (fp64 pipe),
LDS/STS (shared memory), LDG/STG, BAR, MUFU.      python profiles/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "juliaraytracingsw_b200", "lib", "libswrt.so")
KEYS = ["UTMALDG", "UBLKCP", "SYNCS", "LDGSTS", "SHFL", "DFMA", "DADD", "DMUL", "MUFU", "LDS", "STS", "LDG", "STG", "LDL", "STL", "BAR", "ATOM", "RED", "NANOSLEEP"]
proc = subprocess.Popen(["cuobjdump", "-sass", so], stdout=subprocess.PIPE, text=True)
demangle = subprocess.Popen(["c++filt"], stdin=subprocess.PIPE, stdout=subprocess.PIPE, text=True)
hist, name, order = {}, None, []
for line in proc.stdout:
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = m.group(1)
        hist[name] = collections.Counter()
        order.append(name)
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and name:
        hist[name][m.group(1)] += 1
        hist[name]["_total"] += 1
names, _ = demangle.communicate("\n".join(order))
names = names.splitlines()


def short(n):
    n = re.sub(r"\(.*", "", n)                  # drop the argument list
    n = n.replace("swrt::", "").replace("void ", "")
    return n[:110]


total = collections.Counter()
rows = []
for raw, dn in zip(order, names):
    h = hist[raw]
    for k in KEYS:
        total[k] += h[k]
    rows.append((short(dn), h))
print(f"# {os.path.relpath(so, ROOT)}: {len(rows)} kernels; totals: " + ", ".join(f"{k} {total[k]}" for k in KEYS if total[k]))
print("# kernel | instructions | " + " ".join(KEYS))
for n, h in sorted(rows, key=lambda r: r[0]):
    if not (h["UTMALDG"] or h["SHFL"] or h["LDGSTS"] or h["_total"] > 3000 or "raytrace" in n or "team" in n):
        continue                                  # keep the file readable: the large and the interesting kernels
    print(f"{n} | {h['_total']} | " + " ".join(str(h[k]) for k in KEYS))
