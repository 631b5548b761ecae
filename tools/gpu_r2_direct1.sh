#!/bin/bash
# Round 2, first GPU call of the direct path (B2R_OPT_DIRECT_MAX_AREA): the -m gpu suite, then A/B bench lines - direct path
# off / thresholds / setup-kernel occupancy variants - on BASELINE configs 3, 1, 2 and 5, then launch list + --set full of config 3.
#   gpurun --timeout 1500 -- bash tools/gpu_r2_direct1.sh [tag]
set -u
TAG=${1:-d1}
O=gpurun_out/direct_$TAG; mkdir -p $O
python -c "import torch" 2>/dev/null
if [ "${2:-tests}" = tests ]; then timeout 900 python -m pytest tests -m gpu -q -x -n 4 > $O/pytest_gpu.log 2>&1; echo "exit $?" >> $O/pytest_gpu.log; tail -n 5 $O/pytest_gpu.log; fi
B="timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-fill"
for a in 0 16 64; do
	B2R_DIRECT_MAX_AREA=$a $B > $O/bench_spheres_a$a.json 2> $O/bench_spheres_a$a.err
done
for v in none; do
	[ -f build/variants/libb2r_$v.so ] && B2R_LIBRARY=$PWD/build/variants/libb2r_$v.so $B > $O/bench_spheres_$v.json 2> $O/bench_spheres_$v.err
done
for w in vehicle480 vehicle1080 views172; do
	for a in 0 16 64; do
		B2R_DIRECT_MAX_AREA=$a $B --workload $w > $O/bench_${w}_a$a.json 2> $O/bench_${w}_a$a.err
	done
done
timeout 300 python bench.py --steps 40 --no-cpu-baseline --workload fill64 > $O/bench_fill64.json 2> $O/bench_fill64.err
for f in $O/bench_*.json; do echo "$f: $(python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d["value"],1), d["unit"], "| e2e", round(d["e2e"]["value"],1), "|", {k: round(v*1e3,1) for k,v in d.get("stages_ms",{}).items() if v}, "| crc", d.get("checksum",{}).get("match"), "| pairs", d.get("tile_pairs"), "| direct", d.get("direct_max_area"))
except Exception as e: print("FAILED", e)
PY
)"; done | tee $O/summary.txt
export B2R_OVERLAP_FRONT=0 B2R_OVERLAP_VERTEX=0 B2R_FRAME_RING=0
BN="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-fill"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 20 --csv --log-file $O/launches_spheres10m.csv $BN > $O/launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -s 30 -c 10 -f -o $O/spheres10m $BN > $O/full.log 2>&1
ls -la $O
