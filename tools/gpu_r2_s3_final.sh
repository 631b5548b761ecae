#!/bin/bash
# Round 2, session 3: final single-GPU evidence call of the shipped build - the -m gpu suite, the headline bench line (default
# flags), the same without the raster grid hint, the optional workload lines (BASELINE configs 1, 2, 4, 5), then the launch list
# of one frame and a `--set full` capture of the frame's kernels (config 3).
#   gpurun --timeout 290 -- bash tools/gpu_r2_s3_final.sh
set -u
O=gpurun_out/s3_final; mkdir -p $O
python -c "import torch" 2>/dev/null
timeout 200 python -m pytest tests -m gpu -q -x -n 6 > $O/pytest_gpu.log 2>&1; echo "exit $?" >> $O/pytest_gpu.log
tail -n 3 $O/pytest_gpu.log
timeout 120 python bench.py > $O/bench_default.json 2> $O/bench_default.err
B2R_RASTER_GRID_HINT=0 timeout 60 python bench.py --steps 40 --no-cpu-baseline --no-fill > $O/bench_nohint.json 2> $O/bench_nohint.err
for w in vehicle480 vehicle1080 fill64 views172; do
	timeout 60 python bench.py --workload $w --steps 40 --no-cpu-baseline > $O/bench_$w.json 2> $O/bench_$w.err
done
for f in $O/bench_*.json; do echo "$f: $(python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d["value"],2), d["unit"], "| e2e", round(d["e2e"]["value"],2), "|", {k: round(v*1e3,1) for k,v in d.get("stages_ms",{}).items() if v}, "| crc", d.get("checksum",{}).get("match"), "| fill", d.get("fill_test",{}).get("frac_of_hbm_peak"), "| launches", d.get("gpu_launches_per_step"))
except Exception as e: print("FAILED", e)
PY
)"; done | tee $O/summary.txt
# profiles: serial frames under the profiler (ncu serialises launches anyway)
export B2R_OVERLAP_FRONT=0 B2R_OVERLAP_VERTEX=0 B2R_FRAME_RING=0
BENCH="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-fill"
timeout 60 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 22 --csv --log-file $O/launches_spheres10m.csv $BENCH > $O/launches.log 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -s 30 -c 11 -f -o $O/spheres10m $BENCH > $O/full.log 2>&1
ls -la $O | tail -n 5
