#!/bin/bash
# One gpurun call: full -m gpu suite (default mode), the parity suites with the pipelined front end forced on,
# the default bench line, and A/B bench runs (pipelined front end, stream priority, raster small-triangle threshold).
# Everything lands in gpurun_out/ab/.
set -u
O=gpurun_out/ab; mkdir -p $O
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/timeline.txt; }
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.draw --format=csv > $O/gpu.txt 2>&1
el "import torch"; python -c "import torch; print(torch.cuda.get_device_name(0))" > $O/torch.txt 2>&1
el "pytest default"; timeout 600 python -m pytest tests -m gpu -q -x -n 4 --durations=12 > $O/pytest_default.log 2>&1; echo "exit $?" >> $O/pytest_default.log
el "pytest pipelined"; B2R_OVERLAP_FRONT=1 timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_veneer.py tests/test_gpu_units.py -m gpu -q -n 4 > $O/pytest_pipelined.log 2>&1; echo "exit $?" >> $O/pytest_pipelined.log
el "smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "exit $?" >> $O/smoke.log
el "bench default"; timeout 400 python bench.py > $O/bench_default.json 2> $O/bench_default.err
B="timeout 200 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-fill"
el "A base";           $B > $O/ab_base.json 2> $O/ab_base.err
el "A pipe";           B2R_OVERLAP_FRONT=1 $B > $O/ab_pipe.json 2> $O/ab_pipe.err
el "A pipe prio";      B2R_OVERLAP_FRONT=1 B2R_FRONT_PRIORITY=1 $B > $O/ab_pipe_prio.json 2> $O/ab_pipe_prio.err
el "A prio";           B2R_FRONT_PRIORITY=1 $B > $O/ab_prio.json 2> $O/ab_prio.err
el "A pipe prio nopdl"; B2R_OVERLAP_FRONT=1 B2R_FRONT_PRIORITY=1 B2R_PDL=0 $B > $O/ab_pipe_prio_nopdl.json 2> $O/ab_pipe_prio_nopdl.err
el "reference arm + reference CUDA path"; timeout 420 python bench.py --impl reference --steps 3 --warmup 1 --ref-cuda 2 --ref-cuda-timeout 300 > $O/bench_reference.json 2> $O/bench_reference.err
el "A small24";        B2R_LIBRARY=$PWD/build/variants/libb2r_small24.so $B > $O/ab_small24.json 2> $O/ab_small24.err
el "A small96";        B2R_LIBRARY=$PWD/build/variants/libb2r_small96.so $B > $O/ab_small96.json 2> $O/ab_small96.err
el "views172 base";    timeout 200 python bench.py --workload views172 --steps 64 > $O/views_base.json 2> $O/views_base.err
el "views172 pipe";    B2R_OVERLAP_FRONT=1 B2R_FRONT_PRIORITY=1 timeout 200 python bench.py --workload views172 --steps 64 > $O/views_pipe.json 2> $O/views_pipe.err
el "done"
for f in $O/ab_*.json $O/views_*.json; do echo "$f: $(python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(round(d["value"],1), d["unit"], "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1))
except Exception as e: print("FAILED", e)
PY
)"; done | tee $O/summary.txt
tail -3 $O/pytest_default.log $O/pytest_pipelined.log $O/smoke.log
