#!/bin/bash
# Round 2, session 3: single-GPU A/B of the setup-kernel changes (sign-only cull, records written straight to shared memory,
# per-warp statistics) - the in-tree library against build/variants/libb2r_base.so (commit 881f349) and the exact-cull build -
# then the -m gpu suite on the in-tree library.      gpurun --timeout 900 -- bash tools/gpu_r2_s3_ab.sh [tag] [tests|notests]
set -u
O=gpurun_out/s3_${1:-a}; mkdir -p $O
python -c "import torch" 2>/dev/null
B="timeout 200 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-fill"
run() { # name, lib ('' = in-tree), extra args...
	local name=$1 lib=$2; shift 2
	if [ -n "$lib" ]; then B2R_LIBRARY=$PWD/build/variants/libb2r_$lib.so $B "$@" > $O/bench_$name.json 2> $O/bench_$name.err
	else $B "$@" > $O/bench_$name.json 2> $O/bench_$name.err; fi
}
run main_1 ""; run base_1 base; run exactcull_1 exactcull; run main_2 ""; run base_2 base
for w in vehicle1080 views172; do run main_$w "" --workload $w; run base_$w base --workload $w; done
B2R_DIRECT_MAX_AREA=0 run main_a0 ""; B2R_DIRECT_MAX_AREA=0 run base_a0 base
for f in $O/bench_*.json; do echo "$f: $(python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d["value"],1), d["unit"], "| e2e", round(d["e2e"]["value"],1), "|", {k: round(v*1e3,1) for k,v in d.get("stages_ms",{}).items() if v}, "| crc", d.get("checksum",{}).get("match"), d.get("checksum",{}).get("framebuffer_crc32"), "| vis", d.get("visible_triangles"), "| pairs", d.get("tile_pairs"))
except Exception as e: print("FAILED", e)
PY
)"; done | tee $O/summary.txt
if [ "${2:-tests}" = tests ]; then timeout 600 python -m pytest tests -m gpu -q -x -n 4 > $O/pytest_gpu.log 2>&1; echo "exit $?" >> $O/pytest_gpu.log; tail -n 4 $O/pytest_gpu.log; fi
