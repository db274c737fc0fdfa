"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv`) per kernel: launches, total time, share of the step.
usage: launch_summary.py launches.csv > summary.txt"""
import collections
import csv
import re
import sys

OURS = ("msda3d::", "convtc::", "convgen::", "roiattn::", "winattn::", "instnorm::", "tcgemm::", "stemconv::", "fusedln::")
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr, data = rows[0], rows[1:]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
tot = collections.defaultdict(lambda: [0, 0.0])
for r in data:
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "")[:110]
    tot[name][0] += 1
    tot[name][1] += float(r[vi].replace(",", "")) / 1e6
total = sum(v[1] for v in tot.values())
ours = sum(v[1] for k, v in tot.items() if any(o in k for o in OURS))
print(f"# {len(data)} launches, {total:.2f} ms of kernel time (cold-cache, serialised: compare SHARES, not absolutes)")
print(f"# kernels of this library: {ours:.2f} ms = {ours / total:.3f} of the step's kernel time, "
      f"{sum(v[0] for k, v in tot.items() if any(o in k for o in OURS))} launches")
for name, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:60]:
    tag = "*" if any(o in name for o in OURS) else " "
    print(f"{tag} {ms:9.3f} ms  {ms / total:6.3f}  x{n:<4d} {name}")
