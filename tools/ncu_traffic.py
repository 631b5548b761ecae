#!/usr/bin/env python
"""profiles/ncu_traffic.json from `ncu --set full` captures: physical DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum),
duration, SM throughput and threads per instruction of every kernel of one frame, summed per bench.py stage.
usage: tools/ncu_traffic.py spheres10m=path/spheres10m.ncu-rep [fill64=path/fill64.ncu-rep] > profiles/ncu_traffic.json"""
import csv
import json
import subprocess
import sys

STAGE_OF = {"build_constants_kernel": "vertex", "vertex_transform_kernel": "vertex", "frame_clear_kernel": "vertex",
            "triangle_setup_kernel": "setup", "bin_big_count_kernel": "binning", "tile_scan_kernel": "binning", "bin_scatter_kernel": "binning",
            "bin_sort_long_kernel": "binning", "bin_sort_tiles_kernel": "binning", "raster_shade_kernel": "raster_shade", "raster_tail_kernel": "raster_shade", "publish_info_kernel": "raster_shade"}


def num(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return 0.0


out = {"_note": "per stage of one frame: physical DRAM bytes, kernel time, and the issue-slot picture of its dominant kernel; from ncu --set full "
                "(cold caches, serialised launches: the SHARES are what carries over to the timed run, not the absolute times)"}
for arg in sys.argv[1:]:
    name, path = arg.split("=", 1)
    rows = list(csv.reader(subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
    h, units, body = rows[0], rows[1], rows[2:]
    col = {n: i for i, n in enumerate(h)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    stages, seen = {}, set()
    for r in body:
        k = r[col["Kernel Name"]].split("(")[0].split("<")[0].replace("void ", "").replace("b2r::", "")
        st = STAGE_OF.get(k)
        if st is None or (k in seen and k != "raster_shade_kernel"):
            continue           # one frame: the first launch of every kernel
        if k in seen:
            continue
        seen.add(k)
        rd = num(r[col["dram__bytes_read.sum"]]) * scale.get(units[col["dram__bytes_read.sum"]], 1.0)
        wr = num(r[col["dram__bytes_write.sum"]]) * scale.get(units[col["dram__bytes_write.sum"]], 1.0)
        t = num(r[col["gpu__time_duration.sum"]]) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(units[col["gpu__time_duration.sum"]], 1.0)
        s = stages.setdefault(st, {"dram_bytes": 0.0, "time_us": 0.0, "kernels": {}})
        s["dram_bytes"] += rd + wr
        s["time_us"] += t
        s["kernels"][k] = {"time_us": round(t, 2), "dram_bytes": rd + wr, "sm_pct": num(r[col["sm__throughput.avg.pct_of_peak_sustained_elapsed"]]),
                           "dram_pct": num(r[col["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]]),
                           "threads_per_inst": num(r[col["smsp__thread_inst_executed_per_inst_executed.ratio"]]),
                           "issue_active_pct": num(r[col["sm__issue_active.avg.pct_of_peak_sustained_elapsed"]]) if "sm__issue_active.avg.pct_of_peak_sustained_elapsed" in col else None}
    out[name] = stages
print(json.dumps(out, indent=1))
