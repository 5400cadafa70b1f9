#!/usr/bin/env python
"""Per-source-line instruction / stall-sample shares from `ncu --page source --print-source cuda,sass --csv`."""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
only = sys.argv[3] if len(sys.argv) > 3 else ""  # substring of the kernel's function name, e.g. "k_descriptor<(int)0>"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
data = {}
fn_ok = True
for r in rows:
    if r and r[0] == "Function Name":
        fn_ok = only in r[1]
        continue
    if not fn_ok:
        continue
    if r and r[0] == "Line No":
        hdr = r
        ie, smp = hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr is None or len(r) <= ie or not r[0]:
        continue
    try:
        key = (int(r[0]), r[1].strip())
        n, s = int(r[ie]), int(r[smp] or 0)
    except ValueError:
        continue
    a = data.setdefault(key, [0, 0])
    a[0] += n
    a[1] += s
tot = sum(v[0] for v in data.values()) or 1
tots = sum(v[1] for v in data.values()) or 1
print(f"total warp-instructions {tot}, samples {tots}")
for (ln, src), (n, s) in sorted(data.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{100 * n / tot:5.1f}% inst {100 * s / tots:5.1f}% samples  L{ln}: {src[:100]}")
