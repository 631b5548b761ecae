#!/bin/bash
# N-GPU confirmation of the final build: bench.py exactly as the driver launches it (default flags), then with more steps
set -u
N=${1:-8}
O=gpurun_out/r2multig$N; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
port=29970
for steps in 20 100; do
	port=$((port + 1))
	timeout 300 $TR --master-port $port bench.py --gpus $N --steps $steps --warmup 5 > $O/bench_steps$steps.json 2> $O/bench_steps$steps.err
	python - "$O/bench_steps$steps.json" "steps$steps" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("==", sys.argv[2], round(d["value"],1), "fps | e2e", round(d["e2e"]["value"],1), "| crc", d["checksum"]["match"], "| ring", d.get("frame_ring"), "| direct", d.get("direct_max_area"), "| host enqueue ms", round(d["host_enqueue_ms_per_frame"],4), "| launches", d["gpu_launches_per_step"], "| view-parallel", round(d["view_parallel"]["value"],1), "| bands", d["e2e"].get("band_rows_per_rank"))
    for r in d.get("stages_ms_per_rank",[])[:2]: print("     ", r)
except Exception as e: print("   no bench line:", sys.argv[2], e)
PY
done
