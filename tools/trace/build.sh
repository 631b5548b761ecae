#!/usr/bin/env bash
# TEST INFRASTRUCTURE -- builds the GPU-less tracing variant of the product library into build/trace/ (git-ignored):
#   libfake_cudart.so   the recording stand-in for the CUDA runtime (tools/trace/fake_cudart.cpp)
#   libb2r_trace.so     the product's sources, unchanged, compiled with `nvcc -cudart none` and linked against it
#   kernel_params.txt   kernel-parameter layouts from the cubin
# `build.sh --asan` builds the same pair with AddressSanitizer + UndefinedBehaviorSanitizer into build/trace_asan/ (run python under
# LD_PRELOAD=$(g++ -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0): host-side memory errors of the product code.
# `build.sh --tsan` builds them with ThreadSanitizer into build/trace_tsan/ together with tests/trace_scenarios/threads.cpp.
# Use:  B2R_LIBRARY=build/trace/libb2r_trace.so B2R_TRACE_FILE=log.jsonl B2R_TRACE_PARAMS=build/trace/kernel_params.txt python ...
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
OUT="$ROOT/build/trace"
SAN=""; XSAN=""
if [ "${1:-}" = "--tsan" ]; then OUT="$ROOT/build/trace_tsan"; SAN="-g -fsanitize=thread"; XSAN=",-fsanitize=thread"; SANLIB="-ltsan"; fi
if [ "${1:-}" = "--asan" ]; then SANLIB="-lubsan"; OUT="$ROOT/build/trace_asan"; SAN="-g -fsanitize=address"; XSAN=",-fsanitize=address,-fsanitize=undefined,-fno-sanitize-recover=undefined,-fno-omit-frame-pointer"; fi
CSRC="$ROOT/custom_gpgpu_rasterizer_b200/csrc"
mkdir -p "$OUT"
g++ -O1 -std=c++17 -fPIC -shared $SAN -o "$OUT/libfake_cudart.so" "$HERE/fake_cudart.cpp"
if [ ! -f "$OUT/libb2r_trace.so" ] || [ -n "$(find "$CSRC" "$ROOT/include" -newer "$OUT/libb2r_trace.so" \( -name '*.cu' -o -name '*.cuh' -o -name '*.cpp' -o -name '*.hpp' -o -name '*.h' \) -print -quit)" ]; then
  (cd "$CSRC" && nvcc -O1 -std=c++17 -gencode arch=compute_100a,code=sm_100a -fmad=false -Xcompiler -fPIC,-ffp-contract=off$XSAN ${SAN:+-g} -w -cudart none -shared \
      -o "$OUT/libb2r_trace.so" b2r.cu host/b2r_scene.cpp -lz ${SANLIB:-} -L"$OUT" -lfake_cudart -Xlinker -rpath,"$OUT")
fi
if [ "${1:-}" = "--tsan" ]; then
  # the multi-threaded scenario is a C++ program (Python under TSan is noise): four contexts on four threads
  g++ -O1 -g -std=c++17 -fsanitize=thread "$ROOT/tests/trace_scenarios/threads.cpp" -I"$ROOT/include" -L"$OUT" -lb2r_trace -lfake_cudart -Wl,-rpath,"$OUT" -o "$OUT/threads"
fi
python3 "$HERE/kernel_params.py" "$OUT/libb2r_trace.so" "$OUT/kernel_params.txt"
