#!/bin/bash
# Round 2, first single-GPU call: (1) the whole -m gpu suite on the default scheduler, (2) the parity suites again with the
# frame ring (B2R_FRAME_RING=3: every context, single GPU included, then runs enqueueFrameRing), (3) bench lines of both.
#   gpurun --timeout 1500 -- bash tools/gpu_r2_call1.sh
set -u
O=gpurun_out/r2c1; mkdir -p $O
python -c "import torch" 2>/dev/null
timeout 900 python -m pytest tests -m gpu -q -x -n 4 > $O/pytest_default.log 2>&1; echo "exit $?" >> $O/pytest_default.log
B2R_FRAME_RING=3 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_units.py tests/test_gpu_veneer.py tests/test_gpu_zz_fullsize.py -m gpu -q -n 4 > $O/pytest_ring3.log 2>&1; echo "exit $?" >> $O/pytest_ring3.log
B="timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-fill"
$B > $O/bench_default.json 2> $O/bench_default.err
for d in 2 3 4; do B2R_FRAME_RING=$d $B > $O/bench_ring$d.json 2> $O/bench_ring$d.err; done
# BASELINE config 5 (views of the 172-instance scene): shorter frames, more launch-latency-bound - the ring may pay off on one GPU too
timeout 200 python bench.py --workload views172 --steps 64 > $O/views_default.json 2> $O/views_default.err
B2R_FRAME_RING=3 timeout 200 python bench.py --workload views172 --steps 64 > $O/views_ring3.json 2> $O/views_ring3.err
for f in $O/views_*.json; do echo "$f: $(python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print(round(d['value'],1), d['unit'], 'e2e', round(d['e2e']['value'],1))" 2>&1 | tail -1)"; done | tee $O/views_summary.txt
for f in $O/bench_*.json; do echo "$f: $(python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); s=d["stages_ms"]
    print(round(d["value"],1), "fps | e2e", round(d["e2e"]["value"],1), "| " + " ".join(f"{k} {v*1e3:.1f}" for k,v in s.items() if v))
except Exception as e: print("FAILED", e)
PY
)"; done | tee $O/summary.txt
tail -n 3 $O/pytest_default.log $O/pytest_ring3.log
