#!/usr/bin/env python
"""Copy the judged evidence of one gpurun call from gpurun_out/ (scratch) to profiles/ (tracked).

    python tools/collect_profiles.py r01j

JSON lines are stripped of the step banner, .ncu-rep files become summary CSVs (tools/ncu_summary.py), launch
lists and sanitizer logs are copied as they are (sanitizer logs: summary lines only)."""
import glob
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC, DST = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def main():
    tag = sys.argv[1]
    for p in sorted(glob.glob(os.path.join(SRC, tag + "_*"))):
        name = os.path.basename(p)
        if name.endswith(".json"):
            lines = [l for l in open(p) if l.startswith("{")]
            if lines:
                open(os.path.join(DST, name), "w").write(lines[-1])
        elif name.endswith(".ncu-rep"):
            out = os.path.join(DST, name.replace("_prof_", "_ncu_").replace(".ncu-rep", "_summary.csv"))
            subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), p, out], check=True)
        elif name.endswith("_launches.csv"):
            shutil.copy(p, os.path.join(DST, name))
        elif name.endswith("check.log"):
            keep = [l for l in open(p, errors="replace") if "SUMMARY" in l or "sanitize run" in l or "Race reported" in l
                    or "and Read access" in l or "and Write access" in l]
            open(os.path.join(DST, name), "w").write("".join(keep[:40]))
        elif name.endswith("_pytest.log"):
            open(os.path.join(DST, name), "w").write("".join(open(p).readlines()[-3:]))
    print(sorted(os.listdir(DST)))


if __name__ == "__main__":
    main()
