#!/bin/bash
# Round 2, session 3, second single-GPU A/B: empty-tile batches of the raster kernel (in-tree: 16 blocks per work item, a warp
# per block; variants 8 / 32), vertices per CTA of the vertex stage (variants 2 / 8 chunks; in-tree 4), and what the raster
# kernel's surplus CTAs cost (B2R_DEBUG_RASTER_GRID clamps the grid to just above the frame's work items).
set -u
O=gpurun_out/s3_${1:-b}; mkdir -p $O
python -c "import torch" 2>/dev/null
B="timeout 200 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-fill"
run() { local name=$1 lib=$2; shift 2
	if [ -n "$lib" ]; then B2R_LIBRARY=$PWD/build/variants/libb2r_$lib.so $B "$@" > $O/bench_$name.json 2> $O/bench_$name.err
	else $B "$@" > $O/bench_$name.json 2> $O/bench_$name.err; fi
}
run main ""; for v in eb8 eb32 vs8 vs2; do run $v $v; done
B2R_DEBUG_RASTER_GRID=12288 run main_grid12288 ""
run main_vehicle1080 "" --workload vehicle1080
for f in $O/bench_*.json; do echo "$f: $(python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d["value"],1), d["unit"], "| e2e", round(d["e2e"]["value"],1), "|", {k: round(v*1e3,1) for k,v in d.get("stages_ms",{}).items() if v}, "| crc", d.get("checksum",{}).get("match"), d.get("checksum",{}).get("framebuffer_crc32"))
except Exception as e: print("FAILED", e)
PY
)"; done | tee $O/summary.txt
timeout 100 python tools/diag_stage_scaling.py 3 2>&1 | tee $O/stage_scaling.txt
