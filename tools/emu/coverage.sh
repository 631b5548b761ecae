#!/usr/bin/env bash
# TEST INFRASTRUCTURE -- line coverage of the product's KERNEL sources by the emulated scenarios: builds the emulation with
# gcov instrumentation into build/emu_cov, runs a handful of tests/emu_scenarios.py cases and prints, per source file, the
# share of source lines executed in at least one template instantiation (plus the raster kernel's lines that were not).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
OUT="$ROOT/build/emu_cov"
CSRC="$ROOT/custom_gpgpu_rasterizer_b200/csrc"
rm -rf "$OUT"; mkdir -p "$OUT/src"
python3 "$HERE/preprocess.py" "$CSRC" "$OUT/src"
g++ -O1 -std=c++17 -fPIC -shared -o "$OUT/libemu_runtime.so" "$HERE/emu_runtime.cpp"
INC=(-I/usr/local/cuda/include -I"$CSRC" -I"$ROOT/include")
(cd "$OUT" && g++ -std=c++17 -O0 -g --coverage -fPIC -ffp-contract=off -mfma -w "${INC[@]}" -include "$HERE/emu_cuda.h" -c src/b2r.cpp -o b2r.o)
g++ -std=c++17 -O1 -fPIC -ffp-contract=off -w "${INC[@]}" -c "$CSRC/host/b2r_scene.cpp" -o "$OUT/b2r_scene.o"
g++ -shared --coverage -o "$OUT/libb2r_emu.so" "$OUT/b2r.o" "$OUT/b2r_scene.o" -lz -L"$OUT" -lemu_runtime -Wl,-rpath,"$OUT"
export B2R_LIBRARY="$OUT/libb2r_emu.so"
S="$ROOT/tests/emu_scenarios.py"
for c in soup_textured_linear hostile_textured vehicle_normals quads_linear depth_vis; do python3 "$S" parity $c > /dev/null; done
python3 "$S" schedulers ring3 > /dev/null
python3 "$S" schedulers unsorted > /dev/null
B2R_TRACE_DEVICES=2 python3 "$S" sharded 2 ring3sliced 7 > /dev/null
B2R_TRACE_DEVICES=2 python3 "$S" sharded 2 default > /dev/null
B2R_TRACE_DEVICES=3 python3 "$S" replicated 3 > /dev/null
(cd "$OUT" && gcov -o . src/b2r.cpp > /dev/null 2>&1 && python3 - <<'PY'
import collections, re
for f in ("raster_kernel.cuh", "geometry_kernels.cuh", "device_types.cuh", "b2r.cpp"):
    ex, seen = collections.defaultdict(int), set()
    for l in open(f + ".gcov", errors="replace"):
        m = re.match(r"\s*([0-9#=\-\*]+)\*?:\s*(\d+):", l)
        if not m or m.group(1) == "-" or int(m.group(2)) == 0:
            continue
        seen.add(int(m.group(2)))
        if m.group(1) not in ("#####", "====="):
            ex[int(m.group(2))] += 1
    hit = sum(1 for n in seen if ex[n])
    print(f"{f}: {hit}/{len(seen)} source lines executed = {100 * hit / len(seen):.1f} %")
    if f == "raster_kernel.cuh":
        src = open("src/" + f).read().split("\n")
        for n in sorted(n for n in seen if not ex[n]):
            print(f"    not executed: {n}: {src[n - 1].strip()[:100]}")
PY
)
