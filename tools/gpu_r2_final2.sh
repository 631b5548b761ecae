#!/bin/bash
# Round 2, final single-GPU evidence call (direct path build): the whole -m gpu suite, the headline bench line (default flags),
# the optional workload lines (BASELINE configs 1, 2, 4, 5), the launch list of one frame and `--set full` captures of the frame's
# kernels (config 3) and of the raster kernel on config 4.
#   gpurun --timeout 2400 -- bash tools/gpu_r2_final1.sh [tag]
set -u
TAG=${1:-r02}
O=gpurun_out/final2_$TAG; mkdir -p $O
python -c "import torch" 2>/dev/null
timeout 1500 python -m pytest tests -m gpu -q -x -n 4 > $O/pytest_gpu.log 2>&1; echo "exit $?" >> $O/pytest_gpu.log
tail -n 3 $O/pytest_gpu.log
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err
timeout 300 python bench.py --impl reference --ref-cuda 0 > $O/bench_reference.json 2> $O/bench_reference.err
for w in vehicle480 vehicle1080 fill64 views172; do
	timeout 300 python bench.py --workload $w --steps 40 --no-cpu-baseline > $O/bench_$w.json 2> $O/bench_$w.err
done
for f in $O/bench_*.json; do echo "$f: $(python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d["value"],2), d["unit"], "| e2e", round(d["e2e"]["value"],2), "|", {k: round(v*1e3,1) for k,v in d.get("stages_ms",{}).items() if v}, "| crc", d.get("checksum",{}).get("match"), "| fill", d.get("fill_test",{}).get("frac_of_hbm_peak"))
except Exception as e: print("FAILED", e)
PY
)"; done | tee $O/summary.txt
# profiles: serial frames under the profiler (ncu serialises launches anyway)
export B2R_OVERLAP_FRONT=0 B2R_OVERLAP_VERTEX=0 B2R_FRAME_RING=0
BENCH="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-fill"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 20 --csv --log-file $O/launches_spheres10m.csv $BENCH > $O/launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -s 30 -c 10 -f -o $O/spheres10m $BENCH > $O/full.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:raster_shade -s 4 -c 1 -f -o $O/fill64 $BENCH --workload fill64 > $O/fill.log 2>&1
ls -la $O
