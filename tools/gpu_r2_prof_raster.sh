#!/bin/bash
# one `--set full` capture (with source) of the raster kernel of BASELINE config 3, plus the launch list of a frame
set -u
TAG=${1:-r02a}
O=gpurun_out/prof_$TAG; mkdir -p $O
python -c "import torch" 2>/dev/null
BENCH="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-fill"
export B2R_OVERLAP_FRONT=0 B2R_OVERLAP_VERTEX=0
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 24 --csv --log-file $O/launches_spheres10m.csv $BENCH > $O/launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:raster_shade -s 4 -c 1 -f -o $O/raster $BENCH > $O/full.log 2>&1
ls -la $O
