#!/bin/bash
# last single-GPU check of the final build: smoke(), the -m gpu suite, the default bench line and the reference arm
set -u
O=gpurun_out/final3; mkdir -p $O
python -c "import torch" 2>/dev/null
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "exit $?" >> $O/smoke.log; tail -n 2 $O/smoke.log
timeout 900 python -m pytest tests -m gpu -q -x -n 4 > $O/pytest_gpu.log 2>&1; echo "exit $?" >> $O/pytest_gpu.log; tail -n 3 $O/pytest_gpu.log
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 --ref-cuda 0 > $O/bench_reference.json 2> $O/bench_reference.err
python - <<'PY'
import json
for f in ("gpurun_out/final3/bench_default.json", "gpurun_out/final3/bench_reference.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"], 2), d["unit"], "| e2e", round(d["e2e"]["value"], 2), "| roofline", {k: d.get("roofline", {}).get(k) for k in ("kernel", "frac", "physical_frac", "bound")}, "| crc", d.get("checksum", {}).get("match"), "| clocks", d.get("clocks"))
    except Exception as e:
        print(f, "FAILED", e)
PY
