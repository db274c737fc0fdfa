"""Condense `ncu -i REPORT --page source --csv` (SASS view) into a per-kernel stall picture: warp-stall sampling by reason, by opcode, and the
instructions that collect the most samples.   ncu -i x.ncu-rep --page source --csv > src.csv; python tools/ncu_source_stalls.py src.csv"""
import collections
import csv
import sys


def sections(path):
    rows, secs, cur, i = list(csv.reader(open(path))), [], None, 0
    while i < len(rows):
        r = rows[i]
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "h": rows[i + 1], "data": []}
            secs.append(cur)
            i += 2
            continue
        if cur is not None and len(r) == len(cur["h"]):
            cur["data"].append(r)
        i += 1
    return secs


def main():
    seen = set()
    for sec in sections(sys.argv[1]):
        h, data = sec["h"], sec["data"]
        ix = {n: i for i, n in enumerate(h)}
        stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
        tot = collections.Counter()
        for r in data:
            for s in stalls:
                tot[s] += int(r[ix[s]] or 0)
        T, samp = sum(tot.values()), ix["# Samples"]
        key = (sec["name"], T)
        if key in seen or not T:                     # the report lists a kernel once per view
            continue
        seen.add(key)
        op, cnt = collections.Counter(), collections.Counter()
        for r in data:
            o = [x for x in r[ix["Source"]].split() if not x.startswith("@")]
            o = o[0].split(".")[0] if o else "?"
            op[o] += int(r[samp] or 0)
            cnt[o] += int(r[ix["Instructions Executed"]] or 0)
        print(f"== {sec['name'].split('(const')[0]}")
        print(f"   {len(data)} SASS instructions, {T} warp-stall samples, {sum(cnt.values()) / 1e9:.3f} G warp instructions executed")
        print("   stall reasons: " + ", ".join(f"{s[6:]} {100 * v / T:.1f} %" for s, v in tot.most_common(9)))
        print("   by opcode (share of samples | warp instructions): " +
              ", ".join(f"{o} {100 * v / T:.1f} % | {cnt[o] / 1e6:.0f} M" for o, v in op.most_common(12)))
        print("   instructions with the most samples (share, instruction, two leading reasons):")
        for r in sorted(data, key=lambda r: -int(r[samp] or 0))[:12]:
            st = sorted(((int(r[ix[s]] or 0), s[6:]) for s in stalls), reverse=True)[:2]
            print(f"     {100 * int(r[samp]) / T:5.1f} %  {r[ix['Source']].strip()[:72]:72s} {st[0][1]} {st[0][0]}, {st[1][1]} {st[1][0]}")
        print()


if __name__ == "__main__":
    main()
