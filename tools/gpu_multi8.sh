#!/bin/bash
# N-GPU call (N = $1): bench at N ranks with the serial and the pipelined sharded front end, then the sharded correctness check (pipelined)
set -u
N=${1:-8}
O=gpurun_out/multi$N; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
B2R_OVERLAP_FRONT_SHARDED=1 timeout 200 $TR --master-port 29564 bench.py --gpus $N --steps 60 --warmup 5 > $O/bench_pipe.json 2> $O/bench_pipe.err
B2R_OVERLAP_FRONT_SHARDED=0 timeout 200 $TR --master-port 29563 bench.py --gpus $N --steps 60 --warmup 5 > $O/bench_serial.json 2> $O/bench_serial.err
B2R_OVERLAP_FRONT_SHARDED=1 timeout 200 $TR --master-port 29562 tools/sort_first_check.py sharded > $O/check_pipe.log 2>&1; echo "exit $?" >> $O/check_pipe.log
grep -h "sort-first\|exit" $O/check_pipe.log
for f in $O/bench_*.json; do echo "$f: $(python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print(round(d['value'],1),'fps e2e',round(d['e2e']['value'],1), 'view_parallel', round(d['view_parallel']['value'],1)); print(d.get('stages_ms_per_rank',[None])[0])" 2>&1 | tail -2)"; done
