#!/usr/bin/env bash
# A/B builds of libb2r.so for a GPU call (selected with B2R_LIBRARY=$PWD/build/variants/<name>.so; build/ is git-ignored but
# travels to the GPU box).  Run here before `gpurun`.  Each variant is the product source with one -D override.
set -euo pipefail
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
OUT="$ROOT/build/variants"; mkdir -p "$OUT"
cd "$ROOT/custom_gpgpu_rasterizer_b200/csrc"
build() { # name, extra nvcc flags...
  local name=$1; shift
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC,-ffp-contract=off,-Wall "$@" -shared -o "$OUT/libb2r_$name.so" b2r.cu host/b2r_scene.cpp -lz
  echo "built $OUT/libb2r_$name.so"
}
# round 2, f2: the 8x4-texel tiled texture layout (then B2R_TEX_TILED=1 selects it at run time; tools/gpu_r2_tex.sh)
build textiled -DB2R_TEX_TILED_LAYOUT
