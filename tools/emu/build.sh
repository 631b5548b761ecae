#!/usr/bin/env bash
# TEST INFRASTRUCTURE -- builds build/emu/libb2r_emu.so: the product's kernels and host code compiled by the HOST compiler on
# top of tools/emu/emu_cuda.h (CUDA threads as fibers) and the queued host-memory runtime tools/emu/emu_runtime.cpp.  Kernels really run, on
# the CPU, one block at a time - for functional checks of code paths no GPU was available for.  Not a product path: the
# product library has no CPU fallback.   Use:  B2R_LIBRARY=build/emu/libb2r_emu.so python ...
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
OUT="$ROOT/build/emu"
SAN=""
# --asan: the same with AddressSanitizer into build/emu_asan - "device" allocations are heap blocks and __shared__ arrays are
# statics, so a kernel that reads or writes out of bounds is reported (run python under LD_PRELOAD=$(g++ -print-file-name=libasan.so)
# ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0)
if [ "${1:-}" = "--asan" ]; then OUT="$ROOT/build/emu_asan"; SAN="-g -fsanitize=address -fno-omit-frame-pointer"; fi
CSRC="$ROOT/custom_gpgpu_rasterizer_b200/csrc"
CUDAINC="${CUDAINC:-/usr/local/cuda/include}"
mkdir -p "$OUT/src"
python3 "$HERE/preprocess.py" "$CSRC" "$OUT/src"
g++ -O1 -g -std=c++17 -fPIC -shared $SAN -o "$OUT/libemu_runtime.so" "$HERE/emu_runtime.cpp"
FLAGS=(${B2R_EMU_EXTRA:-} -std=c++17 ${SAN:--O2} ${SAN:+-O1} -fPIC -ffp-contract=off -mfma -fno-fast-math -w -I"$CUDAINC" -I"$CSRC" -I"$ROOT/include")
g++ "${FLAGS[@]}" -include "$HERE/emu_cuda.h" -c "$OUT/src/b2r.cpp" -o "$OUT/b2r.o"
g++ "${FLAGS[@]}" -c "$CSRC/host/b2r_scene.cpp" -o "$OUT/b2r_scene.o"
g++ -shared $SAN -o "$OUT/libb2r_emu.so" "$OUT/b2r.o" "$OUT/b2r_scene.o" -lz -L"$OUT" -lemu_runtime -Wl,-rpath,"$OUT"
echo "built $OUT/libb2r_emu.so"
