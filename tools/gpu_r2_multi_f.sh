#!/bin/bash
# N-GPU call (N = $1): frame-ring depth sweep of the sharded front end with the direct path (bench.py --ring D), bit-identity check at the deepest ring
set -u
N=${1:-8}
O=gpurun_out/r2multif$N; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
port=29950
B2R_FRAME_RING=8 B2R_SHARD_SLICED_VERTEX=1 timeout 300 $TR --master-port $port tools/sort_first_check.py sharded > $O/check_ring8_sliced.log 2>&1; echo "exit $?" >> $O/check_ring8_sliced.log
echo "== check ring8 sliced: $(grep -c identical $O/check_ring8_sliced.log) identical, $(grep -c DIFFERENT $O/check_ring8_sliced.log) different, $(tail -n 1 $O/check_ring8_sliced.log)"
for ring in 4 6 8; do
	port=$((port + 1))
	timeout 300 $TR --master-port $port bench.py --gpus $N --steps 200 --warmup 5 --ring $ring > $O/bench_ring$ring.json 2> $O/bench_ring$ring.err
	python - "$O/bench_ring$ring.json" "ring$ring" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("==", sys.argv[2], round(d["value"],1), "fps | e2e", round(d["e2e"]["value"],1), "| crc", d["checksum"]["match"], "| ring", d.get("frame_ring"), "| direct", d.get("direct_max_area"), "| host enqueue ms", round(d["host_enqueue_ms_per_frame"],4), "| launches", d["gpu_launches_per_step"], "| view-parallel", round(d["view_parallel"]["value"],1))
    for r in d.get("stages_ms_per_rank",[])[:2]: print("     ", r)
except Exception as e: print("   no bench line:", sys.argv[2], e)
PY
done
