#!/bin/bash
# f2 (texture layout): A/B of the 8x4-texel tiled layout against the row-major layout on the textured workloads.
# (As measured for profiles/r02_ab_log.md the layout test was compiled into the default build; it is now the variant build
#  build/variants/libb2r_textiled.so of tools/build_variants.sh: add B2R_LIBRARY=$PWD/build/variants/libb2r_textiled.so.)
set -u
O=gpurun_out/tex_${1:-a}; mkdir -p $O
python -c "import torch" 2>/dev/null
B="timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-fill"
for t in 0 1 0 1; do
  for w in fill64 vehicle1080 views172; do
    B2R_TEX_TILED=$t $B --workload $w > $O/bench_${w}_tiled$t.json 2> $O/bench_${w}_tiled$t.err
    python - "$O/bench_${w}_tiled$t.json" "$w tiled=$t" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], round(d["value"],1), d["unit"], "| raster us", round(d.get("stages_ms",{}).get("raster_shade",0)*1e3,1), "| crc", d.get("checksum",{}).get("match"))
except Exception as e: print(sys.argv[2], "FAILED", e)
PY
  done
done | tee $O/summary.txt
B2R_TEX_TILED=1 timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -n 4 -k "textured or linear or vehicle or quads" > $O/pytest_tiled.log 2>&1 <<< "" ; tail -n 2 $O/pytest_tiled.log
