#!/bin/bash
# quick single-GPU A/B of library variants (tools/build_variants.sh): headline bench line per variant
set -u
O=gpurun_out/ab_${1:-x}; mkdir -p $O
python -c "import torch" 2>/dev/null
B="timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline"
$B > $O/bench_default.json 2> $O/bench_default.err
for v in setup2 light5; do
  [ -f build/variants/libb2r_$v.so ] && B2R_LIBRARY=$PWD/build/variants/libb2r_$v.so $B --no-fill > $O/bench_$v.json 2> $O/bench_$v.err
done
for f in $O/bench_*.json; do echo "$f: $(python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); s=d["stages_ms"]
    ft=d.get("fill_test",{})
    print(round(d["value"],1), "fps | e2e", round(d["e2e"]["value"],1), "| nosort", round(d.get("without_in_frame_list_sort",{}).get("value",0),1), "| " + " ".join(f"{k} {v*1e3:.1f}" for k,v in s.items() if v), "| crc", d["checksum"]["match"], "| fallback", d["list_sorter"], "| fill", round(ft.get("frames_per_s",0),1), round(ft.get("frac_of_hbm_peak",0),3))
except Exception as e: print("FAILED", e)
PY
)"; done | tee $O/summary.txt
