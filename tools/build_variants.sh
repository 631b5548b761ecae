#!/usr/bin/env bash
# A/B builds of libb2r.so for a GPU call (selected with B2R_LIBRARY=$PWD/build/variants/<name>.so; build/ is git-ignored but
# travels to the GPU box).  Run here before `gpurun`.  Each variant is the product source with one -D override.
#   otf4: the light raster instantiation that transforms its vertices itself (sliced vertex stage) at 4 resident CTAs per SM
#         (64 registers, no spills) instead of 5 (48 registers, 148 B of spill stores - ptxas -v)
set -euo pipefail
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
OUT="$ROOT/build/variants"; mkdir -p "$OUT"
cd "$ROOT/custom_gpgpu_rasterizer_b200/csrc"
build() { # name, extra nvcc flags...
  local name=$1; shift
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC,-ffp-contract=off,-Wall "$@" -shared -o "$OUT/libb2r_$name.so" b2r.cu host/b2r_scene.cpp -lz
  echo "built $OUT/libb2r_$name.so"
}
build otf4 -DB2R_RASTER_LIGHT_OTF_CTAS=4
# round 2: the persistent raster kernel (warp 0 sorts the staged list while the other warps sweep it)
#   light5: the light raster instantiation at 5 resident CTAs per SM (48 registers) instead of 6 (40 registers, more spills)
build light5 -DB2R_RASTER_LIGHT_CTAS=5
