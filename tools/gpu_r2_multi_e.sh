#!/bin/bash
# N-GPU call (N = $1) for the direct path across ranks (key-buffer blocks pushed to the tile owners): bit-identity checks
# (frame ring 4, plain and sliced vertex stage), then bench lines: direct path on/off x plain/sliced vertex stage.
#   gpurun --gpus N --timeout 1200 -- bash tools/gpu_r2_multi_e.sh N [tag]
set -u
N=${1:-8}
TAG=${2:-e}
O=gpurun_out/r2multi$TAG$N; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
port=29900
for sl in 0 1; do
	port=$((port + 1))
	B2R_FRAME_RING=4 B2R_SHARD_SLICED_VERTEX=$sl timeout 300 $TR --master-port $port tools/sort_first_check.py sharded > $O/check_ring4_sliced$sl.log 2>&1; echo "exit $?" >> $O/check_ring4_sliced$sl.log
	echo "== check ring4 sliced=$sl: $(grep -c identical $O/check_ring4_sliced$sl.log) identical, $(grep -c DIFFERENT $O/check_ring4_sliced$sl.log) different, $(tail -n 1 $O/check_ring4_sliced$sl.log)"
done
for mode in direct_plain direct_sliced binned_plain; do
	port=$((port + 1))
	unset B2R_SHARD_SLICED_VERTEX B2R_DIRECT_MAX_AREA
	[ $mode = direct_sliced ] && export B2R_SHARD_SLICED_VERTEX=1
	[ $mode = binned_plain ] && export B2R_DIRECT_MAX_AREA=0
	timeout 300 $TR --master-port $port bench.py --gpus $N --steps 100 --warmup 5 > $O/bench_$mode.json 2> $O/bench_$mode.err
	python - "$O/bench_$mode.json" "$mode" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("==", sys.argv[2], round(d["value"],1), "fps | e2e", round(d["e2e"]["value"],1), "| crc", d["checksum"]["match"], "| ring", d.get("frame_ring"), "| direct", d.get("direct_max_area"), "| host enqueue ms", round(d["host_enqueue_ms_per_frame"],4), "| launches", d["gpu_launches_per_step"], "| view-parallel", round(d["view_parallel"]["value"],1))
    for r in d.get("stages_ms_per_rank",[])[:2]: print("     ", r)
except Exception as e: print("   no bench line:", sys.argv[2], e)
PY
done
unset B2R_SHARD_SLICED_VERTEX B2R_DIRECT_MAX_AREA
