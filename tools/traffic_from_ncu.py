#!/usr/bin/env python
"""profiles/depth_tile_traffic.json from an `ncu --set full` capture of the configs[2] step: DRAM bytes read + written
by one launch of the depth kernel (what bench.py reports as roofline.traffic).

    python tools/traffic_from_ncu.py gpurun_out/r02x_prof_step.ncu-rep r02x
"""
import csv
import json
import subprocess
import sys


def main():
    rep, tag = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    head, units = rows[0], rows[1]
    kn, rd, wr, tm = (head.index(k) for k in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum",
                                              "gpu__time_duration.sum"))
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for r in rows[2:]:
        if "depth_tile" in r[kn]:
            b = float(r[rd]) * scale[units[rd]] + float(r[wr]) * scale[units[wr]]
            out = {"workload": "c3", "kernel": r[kn].split("(")[0], "dram_bytes_per_launch": b,
                   "dram_bytes_read": float(r[rd]) * scale[units[rd]], "dram_bytes_write": float(r[wr]) * scale[units[wr]],
                   "kernel_time_under_ncu": f"{r[tm]} {units[tm]}", "source": f"profiles/{tag}_ncu_step_summary.csv "
                   "(ncu --set full --clock-control none, GCI_GRAPH=0 bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e)"}
            json.dump(out, open("profiles/depth_tile_traffic.json", "w"), indent=1)
            print(out)
            return
    raise SystemExit("no depth_tile kernel in the capture")


if __name__ == "__main__":
    main()
