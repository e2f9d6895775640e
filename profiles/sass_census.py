#!/usr/bin/env python
"""SASS census of libpbremap.so: per kernel, how many TMA / mbarrier / bulk-copy / cp.async
instructions the compiled sm_100a code holds (static counts from `cuobjdump -sass`; no GPU needed).

    python profiles/sass_census.py > profiles/r2_sass_census.txt

UTMALDG = TMA tensor-map box load, UTMASTG = TMA tensor-map box store, UBLKCP = 1-D bulk copy
(cp.async.bulk), UTMAPF = TMA prefetch to L2, SYNCS = mbarrier operations, LDGSTS = cp.async
(16-byte global -> shared copies), DADD/DMUL/DFMA = float64 arithmetic, FFMA = float arithmetic.
"""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, "photonbend_b200", "libpbremap.so")
WATCH = ("UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "SYNCS", "LDGSTS", "BAR", "LDS", "STS", "DADD", "DMUL", "DFMA", "FFMA", "MUFU")


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else LIB
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    demangle = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True,
                              text=True).stdout.splitlines()
    names = iter(demangle)
    counts, order, cur = {}, [], None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = next(names, m.group(1))
            cur = cur.replace("(int)", "").replace("(bool)", "")
            cur = re.sub(r"\([^()]*\)$", "", cur).replace("void pb::", "").replace("pb::", "")
            counts[cur] = collections.Counter()
            order.append(cur)
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            op = m.group(1)
            counts[cur]["total"] += 1
            for w in WATCH:
                if op == w or op.startswith(w + "."):
                    counts[cur][w] += 1
    print(f"# {os.path.relpath(lib, REPO)}: static SASS instruction counts per kernel (cuobjdump -sass, sm_100a)")
    print(f"{'kernel':58s} {'total':>6s} " + " ".join(f"{w:>7s}" for w in WATCH))
    tot = collections.Counter()
    for k in order:
        c = counts[k]
        tot.update(c)
        print(f"{k[:58]:58s} {c['total']:6d} " + " ".join(f"{c[w]:7d}" for w in WATCH))
    print(f"{'ALL KERNELS':58s} {tot['total']:6d} " + " ".join(f"{tot[w]:7d}" for w in WATCH))


if __name__ == "__main__":
    main()
