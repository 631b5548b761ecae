#!/bin/bash
# quick A/B call: parity suites (4 workers) + bench stage times for the main build and the variants in build/variants
set -u
O=gpurun_out/ab2; mkdir -p $O
python -c "import torch" 2>/dev/null
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_units.py tests/test_gpu_veneer.py -m gpu -q -x -n 4 > $O/pytest.log 2>&1; echo "exit $?" >> $O/pytest.log
B="timeout 200 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-fill"
$B > $O/main.json 2> $O/main.err
B2R_OVERLAP_FRONT=0 $B > $O/main_serial.json 2> $O/main_serial.err
for v in build/variants/*.so; do n=$(basename $v .so); B2R_LIBRARY=$PWD/$v $B > $O/$n.json 2> $O/$n.err; done
for f in $O/*.json; do echo "$f: $(python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); s=d["stages_ms"]
    print(round(d["value"],1), "fps | " + " ".join(f"{k} {v*1e3:.1f}" for k,v in s.items() if v), "| unsorted", round(d.get("without_in_frame_list_sort",{}).get("value",0),1), d.get("list_sorter"))
except Exception as e: print("FAILED", e)
PY
)"; done | tee $O/summary.txt
tail -n 3 $O/pytest.log
