#!/bin/bash
# Round 2, session 3: N-GPU check of the final build (bench.py as the driver launches it) and A/B of the OTF raster
# instantiation's occupancy (build/variants/libb2r_otf{4,5}.so: 4 / 5 CTAs per SM instead of 6 - fewer spills).
#   gpurun --gpus N --timeout 300 -- bash tools/gpu_r2_s3_multi.sh N [variants...]
set -u
N=${1:-4}; shift
O=gpurun_out/s3_multi$N; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
port=29980
for v in main "$@"; do
	port=$((port + 1))
	unset B2R_LIBRARY B2R_RASTER_GRID_HINT
	if [ $v = nohint ]; then export B2R_RASTER_GRID_HINT=0; elif [ $v != main ]; then export B2R_LIBRARY=$PWD/build/variants/libb2r_$v.so; fi
	timeout 120 $TR --master-port $port bench.py --gpus $N --steps 40 --warmup 5 --no-cpu-baseline --no-fill > $O/bench_$v.json 2> $O/bench_$v.err
	python - "$O/bench_$v.json" "$v" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("==", sys.argv[2], round(d["value"],1), "fps | e2e", round(d["e2e"]["value"],1), "| crc", d["checksum"]["match"], "| ring", d.get("frame_ring"), "| launches", d["gpu_launches_per_step"], "| view-parallel", round(d.get("view_parallel",{}).get("value",0),1))
    for r in d.get("stages_ms_per_rank",[])[:2]: print("     ", r)
except Exception as e: print("   no bench line:", sys.argv[2], e)
PY
done
