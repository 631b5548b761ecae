#!/bin/bash
# 2-GPU call: sharded sort-first check (serial and pipelined front end) + bench at N=2 for both
set -u
O=gpurun_out/multi2; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29551 tools/sort_first_check.py sharded > $O/check_serial.log 2>&1; echo "exit $?" >> $O/check_serial.log
B2R_OVERLAP_FRONT_SHARDED=1 timeout 300 $TR --master-port 29552 tools/sort_first_check.py sharded > $O/check_pipe.log 2>&1; echo "exit $?" >> $O/check_pipe.log
timeout 300 $TR --master-port 29553 bench.py --gpus 2 --steps 40 --warmup 5 > $O/bench_serial.json 2> $O/bench_serial.err
B2R_OVERLAP_FRONT_SHARDED=1 timeout 300 $TR --master-port 29554 bench.py --gpus 2 --steps 40 --warmup 5 > $O/bench_pipe.json 2> $O/bench_pipe.err
grep -h "sort-first\|exit" $O/check_serial.log $O/check_pipe.log
for f in $O/bench_*.json; do echo "$f: $(python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print(round(d['value'],1),'fps e2e',round(d['e2e']['value'],1), 'view_parallel', round(d['view_parallel']['value'],1))" 2>&1 | tail -1)"; done
