#!/usr/bin/env bash
# TEST INFRASTRUCTURE -- builds the reference's OWN CPU renderer (Elite::Renderer, CustomCUDARasterizer/ERenderer.cpp)
# into oracle/_ref/libere_ref.so, so that the restatement oracle/erenderer_oracle.cpp can be pinned against it and
# bench.py can time the real thing as the CPU baseline.
#
# ERenderer.cpp is excluded from the reference's project (vcxproj:126-155) because it predates the current Mesh /
# SceneManager interfaces (GetVertexAmount, GetIndexAmount, GetTextures, GetWorldMatrix, IsDepthColour).  The file
# itself - every line of arithmetic on the CPU path - is compiled UNMODIFIED, with the reference's own Camera.cpp,
# Light.cpp, DirectionalLight.cpp, SceneGraph.cpp, Texture.cpp and math / vertex / texture-sampler headers, from a
# BUILD-TIME scratch copy (oracle/_ref/src_ere, git-ignored: reference sources never enter this repository).  What is
# ours: the three headers of oracle/ere_shim/ that restore the interfaces the file was written against (copied over
# the scratch copies of PCH.h, Mesh.h, SceneManager.h under those names), the SDL stand-ins (sdl_stub.cpp, sdl_stub_ere.cpp) and the glue
# (ere_ref_driver.cpp).  One leniency patch, as in build_ref.sh: std::powf -> powf (RGBColor.h:126-128; GCC's <cmath>
# has no std::powf).  Built like the restatement: g++ -O2 -ffp-contract=off (no implicit FMA, IEEE single precision).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${REF:-/root/reference/CustomCUDARasterizer}"
SDLINC="${SDLINC:-/root/reference/3rdParty/include}"
CUDAINC="${CUDAINC:-/usr/local/cuda/include}"
OUT="$HERE/_ref"
SRC="$OUT/src_ere"
CXX="${CXX:-g++}"

if [ ! -d "$REF" ]; then
  echo "build_ref_erenderer: $REF not present (GPU box?) - keeping prebuilt $OUT/libere_ref.so" >&2
  exit 0
fi

rm -rf "$SRC"
mkdir -p "$SRC"
cp "$REF"/*.h "$SRC"/
for f in ERenderer Camera Light DirectionalLight SceneGraph Texture; do cp "$REF/$f.cpp" "$SRC"/; done
# our reconstruction of the interfaces ERenderer.cpp predates replaces the current headers of the same name
cp "$HERE"/ere_shim/erenderer_prelude.h "$SRC"/PCH.h
cp "$HERE"/ere_shim/legacy_mesh_interface.h "$SRC"/Mesh.h
cp "$HERE"/ere_shim/legacy_scene_manager_interface.h "$SRC"/SceneManager.h
sed -i 's/std::powf/powf/g' "$SRC/RGBColor.h"

FLAGS=(-std=c++17 -O2 -fPIC -ffp-contract=off -fno-fast-math -w
       -I"$SRC" -I"$SDLINC/sdl2-2.0.9" -I"$SDLINC/sdl2_image-2.0.5" -I"$CUDAINC")
"$CXX" "${FLAGS[@]}" -shared -o "$OUT/libere_ref.so" \
  "$SRC"/ERenderer.cpp "$SRC"/Camera.cpp "$SRC"/Light.cpp "$SRC"/DirectionalLight.cpp "$SRC"/SceneGraph.cpp "$SRC"/Texture.cpp \
  "$HERE"/ere_ref_driver.cpp "$HERE"/sdl_stub.cpp "$HERE"/sdl_stub_ere.cpp
echo "build_ref_erenderer: built $OUT/libere_ref.so"
