#!/usr/bin/env python
"""Count the SASS mnemonics that show how each kernel of libgci_cuda.so moves data (no GPU needed):
UBLKCP = cp.async.bulk (TMA bulk copy), SYNCS = mbarrier ops, REDUX = warp reductions, ATOMS / ATOMG / RED = atomics,
LDS / STS = shared memory, LDG / STG .128 = 16-byte global accesses, SHFL / VOTE = warp exchange.

    python tools/sass_evidence.py > profiles/sass_mnemonics.txt
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["UBLKCP", "SYNCS", "FENCE", "REDUX", "ATOMS", "ATOMG", "RED", "LDS", "LDS.128", "STS", "LDG", "LDG.128", "STG",
        "STG.128", "SHFL", "VOTE", "BAR"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "gci_b200", "libgci_cuda.so")], capture_output=True,
                          text=True, check=True).stdout
    demangle = {}
    per = collections.OrderedDict()
    fn = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            per[fn] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m or fn is None:
            continue
        op = m.group(1)
        base = op.split(".")[0]
        per[fn]["total"] += 1
        if base in KEYS:
            per[fn][base] += 1
        if base in ("LDS", "LDG", "STG") and ".128" in op:
            per[fn][base + ".128"] += 1
    names = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
    for mangled, nice in zip(per, names):
        demangle[mangled] = re.sub(r"\(.*", "", nice)
    print(f"{'kernel':44s} {'instr':>6s} " + " ".join(f"{k:>7s}" for k in KEYS))
    for fn, c in sorted(per.items(), key=lambda kv: demangle[kv[0]]):
        print(f"{demangle[fn][:44]:44s} {c['total']:6d} " + " ".join(f"{c[k]:7d}" for k in KEYS))


if __name__ == "__main__":
    main()
