#!/usr/bin/env python
"""Where a kernel's instructions go: groups the SASS of an .ncu-rep (captured with --import-source on / -lineinfo) into
runs of consecutive instructions with similar execution counts and prints each run's share of the executed warp
instructions and of the stall samples.   python tools/ncu_regions.py gpurun_out/x.ncu-rep [--sass A:B]"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    isrc, iex, ismp, ith = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Avg. Threads Executed")
    data = rows[2:]
    if "--sass" in sys.argv:
        a, b = map(int, sys.argv[sys.argv.index("--sass") + 1].split(":"))
        for k in range(a, b):
            r = data[k]
            print(k, r[isrc].strip()[:90].ljust(90), r[iex].rjust(10), r[ismp].rjust(6), r[ith])
        return
    tot = sum(int(r[iex]) for r in data)
    ts = sum(int(r[ismp]) for r in data)
    thr = sum(int(r[iex]) * float(r[ith]) for r in data)
    print(f"{rows[0][1][:80]}: {tot / 1e6:.2f} M warp instructions, {len(data)} SASS lines, {thr / tot:.1f} threads per instruction")
    regions, cur = [], None
    for k, r in enumerate(data):
        c = int(r[iex])
        if cur and (0.7 * cur["c"] <= c <= 1.4 * cur["c"] or abs(c - cur["c"]) < 0.0002 * tot):
            cur["n"] += 1
            cur["tot"] += c
            cur["smp"] += int(r[ismp])
            cur["end"] = k
        else:
            cur = {"start": k, "end": k, "c": c, "n": 1, "tot": c, "smp": int(r[ismp])}
            regions.append(cur)
    for g in regions:
        if g["tot"] > 0.004 * tot:
            print(f"{g['start']:5d}-{g['end']:5d} n={g['n']:4d} exec/inst={g['c']:9d} total={g['tot'] / 1e6:7.2f}M ({100 * g['tot'] / tot:4.1f}%) "
                  f"samples {100 * g['smp'] / max(ts, 1):4.1f}%  first: {data[g['start']][isrc].strip()[:50]}")


if __name__ == "__main__":
    main()
