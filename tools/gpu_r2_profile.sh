#!/bin/bash
# Round 2 profile call (one GPU): launch list of the frame (per-launch gpu__time_duration), one `--set full` capture of a whole
# frame of BASELINE config 3 and of the raster kernel on config 4, summarised into markdown (copy the .md / .csv into profiles/).
#   gpurun --timeout 1200 -- bash tools/gpu_r2_profile.sh [tag]
set -u
TAG=${1:-r02}
O=gpurun_out/prof_$TAG; mkdir -p $O
python -c "import torch" 2>/dev/null
BENCH="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-fill"
# serial frames under the profiler (ncu serialises launches anyway); skip the scene-load kernels and two warm-up frames
export B2R_OVERLAP_FRONT=0 B2R_OVERLAP_VERTEX=0
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 26 --csv --log-file $O/launches_spheres10m.csv $BENCH > $O/launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -s 30 -c 13 -f -o $O/spheres10m $BENCH > $O/full.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:raster_shade -s 4 -c 1 -f -o $O/fill64 $BENCH --workload fill64 > $O/fill.log 2>&1
for r in spheres10m fill64; do
	[ -f $O/$r.ncu-rep ] && python tools/profile_summary.py $O/$r.ncu-rep "$TAG: $r" > $O/${TAG}_$r.md 2>> $O/summary.err
done
python - "$O/launches_spheres10m.csv" <<'PY' | tee $O/launch_table.txt
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]; kn, mv = h.index("Kernel Name"), h.index("Metric Value")
for r in rows[hi + 1:]:
    if len(r) > mv and r[mv]:
        print(f"{r[kn].split('(')[0][-44:]:46s} {float(r[mv].replace(',', '')) / 1000:8.1f} us")
PY
ls -la $O
