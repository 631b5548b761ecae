#!/bin/bash
# Round 2, N-GPU call (N = $1): sharded sort-first with the default three-stage pipeline, with the frame ring, and with the
# frame ring + sliced vertex stage.  For every mode: the bit-identity check against one GPU (tools/sort_first_check.py:
# composed frame, concatenated lists, 10 frames in flight with a moving camera), then the bench line.
#   gpurun --gpus 2 --timeout 900 -- bash tools/gpu_r2_multi.sh 2
set -u
N=${1:-2}
O=gpurun_out/r2multi$N; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
port=29600
run() {   # name, env...
	local name=$1; shift
	port=$((port + 1))
	env "$@" timeout 240 $TR --master-port $port tools/sort_first_check.py sharded > $O/check_$name.log 2>&1; echo "exit $?" >> $O/check_$name.log
	if grep -q "^exit 0" $O/check_$name.log; then
		port=$((port + 1))
		env "$@" timeout 240 $TR --master-port $port bench.py --gpus $N --steps 60 --warmup 5 > $O/bench_$name.json 2> $O/bench_$name.err
	else
		echo "check failed: bench skipped" > $O/bench_$name.err
	fi
	echo "== $name: $(grep -c identical $O/check_$name.log) identical, $(grep -c DIFFERENT $O/check_$name.log) different, $(tail -n 1 $O/check_$name.log)"
	python - "$O/bench_$name.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("  ", round(d["value"],1), "fps | e2e", round(d["e2e"]["value"],1), "| crc", d["checksum"]["match"], "| view-parallel", round(d["view_parallel"]["value"],1), "|", d.get("stages_ms_per_rank",[None])[0])
except Exception as e: print("   no bench line:", e)
PY
}
run default B2R_FRAME_RING=0
run ring3 B2R_FRAME_RING=3
run ring4 B2R_FRAME_RING=4
run ring3_sliced B2R_FRAME_RING=3 B2R_SHARD_SLICED_VERTEX=1
# the end-to-end leg through rank 0 alone (colours over NVLink, one PCIe link) next to the distributed Present of the lines above
port=$((port + 1))
B2R_FRAME_RING=3 timeout 240 $TR --master-port $port bench.py --gpus $N --steps 60 --warmup 5 --present central > $O/bench_ring3_central.json 2> $O/bench_ring3_central.err
python - "$O/bench_ring3_central.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print("== ring3, central Present: e2e", round(d["e2e"]["value"],1), "crc", d["checksum"]["match"])
except Exception as e: print("   no bench line:", e)
PY
port=$((port + 1))
B2R_FRAME_RING=3 timeout 240 $TR --master-port $port bench.py --gpus $N --steps 60 --warmup 5 --present tiles > $O/bench_ring3_tiles.json 2> $O/bench_ring3_tiles.err
python - "$O/bench_ring3_tiles.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print("== ring3, Present by tiles: e2e", round(d["e2e"]["value"],1), "crc", d["checksum"]["match"])
except Exception as e: print("   no bench line:", e)
PY
# BASELINE config 5 at N ranks: view-parallel (no communication) and sort-first per view (sharded front end, moving camera)
for m in parallel sortfirst; do
	port=$((port + 1))
	timeout 240 $TR --master-port $port bench.py --gpus $N --workload views172 --views-mode $m --steps 64 > $O/views_$m.json 2> $O/views_$m.err
	echo "== views172 $m: $(python -c "
import json
d=json.loads(open('$O/views_$m.json').read().strip().splitlines()[-1]); print(round(d['value'],1), d['unit'], 'e2e', round(d['e2e']['value'],1))" 2>&1 | tail -1)"
done
