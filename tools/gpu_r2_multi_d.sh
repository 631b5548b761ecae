#!/bin/bash
# N-GPU call (N = $1): bit-identity check, bench lines (frame ring 4; plain / sliced vertex stage / strictly chained raster
# kernels), the host-bandwidth probe of the distributed Present.
set -u
N=${1:-8}
O=gpurun_out/r2multid$N; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
port=29900
B2R_FRAME_RING=4 B2R_SHARD_SLICED_VERTEX=1 timeout 240 $TR --master-port $port tools/sort_first_check.py sharded > $O/check_ring4_sliced.log 2>&1; echo "exit $?" >> $O/check_ring4_sliced.log
echo "== check ring4 sliced: $(grep -c identical $O/check_ring4_sliced.log) identical, $(grep -c DIFFERENT $O/check_ring4_sliced.log) different, $(tail -n 1 $O/check_ring4_sliced.log)"
for mode in plain sliced strict; do
	port=$((port + 1))
	unset B2R_SHARD_SLICED_VERTEX B2R_RING_RELAXED
	[ $mode = sliced ] && export B2R_SHARD_SLICED_VERTEX=1
	[ $mode = strict ] && export B2R_RING_RELAXED=0
	timeout 300 $TR --master-port $port bench.py --gpus $N --steps 100 --warmup 5 > $O/bench_$mode.json 2> $O/bench_$mode.err
	python - "$O/bench_$mode.json" "$mode" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("==", sys.argv[2], round(d["value"],1), "fps | e2e", round(d["e2e"]["value"],1), "| crc", d["checksum"]["match"], "| ring", d.get("frame_ring"), "| host enqueue ms", round(d["host_enqueue_ms_per_frame"],4), "| launches", d["gpu_launches_per_step"], "| view-parallel", round(d["view_parallel"]["value"],1))
    for r in d.get("stages_ms_per_rank",[])[:2]: print("     ", r)
except Exception as e: print("   no bench line:", sys.argv[2], e)
PY
done
unset B2R_SHARD_SLICED_VERTEX B2R_RING_RELAXED
port=$((port + 1))
timeout 200 $TR --master-port $port tools/pcie_probe_multi.py > $O/pcie_probe.txt 2> $O/pcie_probe.err; cat $O/pcie_probe.txt
(nvidia-smi topo -m; lscpu | grep -i -E "numa|socket|model name") > $O/topology.txt 2>&1
