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
    print("  ", round(d["value"],1), "fps | e2e", round(d["e2e"]["value"],1), "| view-parallel", round(d["view_parallel"]["value"],1), "|", d.get("stages_ms_per_rank",[None])[0])
except Exception as e: print("   no bench line:", e)
PY
}
run default B2R_FRAME_RING=0
run ring3 B2R_FRAME_RING=3
run ring4 B2R_FRAME_RING=4
run ring3_sliced B2R_FRAME_RING=3 B2R_SHARD_SLICED_VERTEX=1
run ring4_sliced B2R_FRAME_RING=4 B2R_SHARD_SLICED_VERTEX=1
# programmatic dependent launch parks the next kernel's CTAs on SM slots while the current kernel's tail drains: good for one
# stream, possibly bad when other ring streams have runnable work; more hardware queues avoid false dependencies between streams
run ring3_sliced_nopdl B2R_FRAME_RING=3 B2R_SHARD_SLICED_VERTEX=1 B2R_PDL=0
#run ring3_sliced_conn32 B2R_FRAME_RING=3 B2R_SHARD_SLICED_VERTEX=1 CUDA_DEVICE_MAX_CONNECTIONS=32
# BASELINE config 5 at N ranks: view-parallel (no communication) and sort-first per view (sharded front end, moving camera)
for m in parallel sortfirst; do
	port=$((port + 1))
	timeout 240 $TR --master-port $port bench.py --gpus $N --workload views172 --views-mode $m --steps 64 > $O/views_$m.json 2> $O/views_$m.err
	echo "== views172 $m: $(python -c "
import json
d=json.loads(open('$O/views_$m.json').read().strip().splitlines()[-1]); print(round(d['value'],1), d['unit'], 'e2e', round(d['e2e']['value'],1))" 2>&1 | tail -1)"
done
# the on-the-fly raster variant at 4 resident CTAs per SM (no register spills) - build it first with tools/build_variants.sh
[ -f build/variants/libb2r_otf4.so ] && run ring3_sliced_otf4 B2R_FRAME_RING=3 B2R_SHARD_SLICED_VERTEX=1 B2R_LIBRARY=$PWD/build/variants/libb2r_otf4.so
