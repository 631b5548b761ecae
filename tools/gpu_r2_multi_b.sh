#!/bin/bash
# N-GPU call (N = $1): bit-identity check of the sharded front end, then the bench line (frame ring 4, distributed Present by
# row bands) and the two other Present modes for comparison.
set -u
N=${1:-2}
O=gpurun_out/r2multib$N; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
port=29700
B2R_FRAME_RING=4 timeout 240 $TR --master-port $port tools/sort_first_check.py sharded > $O/check_ring4.log 2>&1; echo "exit $?" >> $O/check_ring4.log
echo "== check ring4: $(grep -c identical $O/check_ring4.log) identical, $(grep -c DIFFERENT $O/check_ring4.log) different, $(tail -n 1 $O/check_ring4.log)"
for mode in band tiles central; do
	port=$((port + 1))
	timeout 300 $TR --master-port $port bench.py --gpus $N --steps 60 --warmup 5 --present $mode > $O/bench_$mode.json 2> $O/bench_$mode.err
	python - "$O/bench_$mode.json" "$mode" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("==", sys.argv[2], round(d["value"],1), "fps | e2e", round(d["e2e"]["value"],1), "| crc", d["checksum"]["match"], "| ring", d.get("frame_ring"), "| view-parallel", round(d["view_parallel"]["value"],1), "|", d.get("stages_ms_per_rank",[None])[0])
except Exception as e: print("   no bench line:", sys.argv[2], e)
PY
done
