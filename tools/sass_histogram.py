#!/usr/bin/env python
"""SASS mnemonic histogram per source file of libmvp_ops.so (no GPU needed):
    python -c "import __graft_entry__ as g; g.build()" && python tools/sass_histogram.py > profiles/rN_sass_histogram.md
Reads the objects build.py leaves under mvp_benchmark_b200/build/ through `cuobjdump -sass`."""
import collections
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NOTE = ["UTCHMMA", "UTCQMMA", "LDTM", "UTCBAR", "UTCATOMSWS", "UTMALDG", "UBLKCP", "SYNCS", "LDGSTS", "CREDUX", "FFMA2", "FADD2", "FMUL2",
        "FMNMX3", "REDG", "ATOMS", "ATOMG", "MATCH", "VOTE", "SHFL", "BAR", "UCGABAR", "ACQBULK", "LDS", "STS", "LDG", "STG", "LDL",
        "STL", "DADD", "MUFU"]


def histogram(obj):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    c = collections.Counter()
    for line in out.splitlines():
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m:
            c[m.group(1)] += 1
    return c


def main():
    objs = sorted(glob.glob(os.path.join(ROOT, "mvp_benchmark_b200", "build", "*.o")))
    if not objs:
        sys.exit("no objects: run __graft_entry__.build() first")
    print("# SASS mnemonic histogram per source file (`cuobjdump -sass` of the objects that make libmvp_ops.so; sm_100a; "
          "`tools/sass_histogram.py`)\n")
    for obj in objs:
        c = histogram(obj)
        name = os.path.basename(obj).replace(".o", ".cu")
        print("## %s: %d instructions\n" % (name, sum(c.values())))
        print("top: " + ", ".join("%s %d" % kv for kv in c.most_common(14)) + "\n")
        print("of note: " + ", ".join("%s %d" % (k, c[k]) for k in NOTE if c[k]) + "\n")


if __name__ == "__main__":
    main()
