#!/bin/bash
# single-GPU A/B of library variants (build/variants/*.so) on BASELINE config 3 (+ the all-binned path and the vehicle scenes), then the -m gpu suite
set -u
O=gpurun_out/ab3_${1:-x}; mkdir -p $O
python -c "import torch" 2>/dev/null
B="timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-fill"
$B > $O/bench_default.json 2> $O/bench_default.err
B2R_DIRECT_MAX_AREA=0 $B > $O/bench_default_a0.json 2> $O/bench_default_a0.err
$B --workload vehicle1080 > $O/bench_default_vehicle1080.json 2> $O/bench_default_vehicle1080.err
$B --workload views172 > $O/bench_default_views172.json 2> $O/bench_default_views172.err
for v in $(ls build/variants/ 2>/dev/null | sed 's/libb2r_//; s/\.so//'); do
  B2R_LIBRARY=$PWD/build/variants/libb2r_$v.so $B > $O/bench_$v.json 2> $O/bench_$v.err
  B2R_DIRECT_MAX_AREA=0 B2R_LIBRARY=$PWD/build/variants/libb2r_$v.so $B > $O/bench_${v}_a0.json 2> $O/bench_${v}_a0.err
  B2R_LIBRARY=$PWD/build/variants/libb2r_$v.so $B --workload vehicle1080 > $O/bench_${v}_vehicle1080.json 2> $O/bench_${v}_vehicle1080.err
done
for f in $O/bench_*.json; do echo "$f: $(python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d["value"],1), d["unit"], "| e2e", round(d["e2e"]["value"],1), "|", {k: round(v*1e3,1) for k,v in d.get("stages_ms",{}).items() if v}, "| crc", d.get("checksum",{}).get("match"), "| direct", d.get("direct_max_area"))
except Exception as e: print("FAILED", e)
PY
)"; done | tee $O/summary.txt
if [ "${2:-tests}" = tests ]; then timeout 900 python -m pytest tests -m gpu -q -x -n 4 > $O/pytest_gpu.log 2>&1; echo "exit $?" >> $O/pytest_gpu.log; tail -n 4 $O/pytest_gpu.log; fi
