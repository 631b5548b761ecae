#!/usr/bin/env bash
# TEST INFRASTRUCTURE -- builds the *unmodified reference renderer* into oracle/_ref/.
#
# What it does: takes the reference's own sources where they lie under
# $REF (default /root/reference/CustomCUDARasterizer), makes a BUILD-TIME scratch
# copy under oracle/_ref/src (git-ignored and gpurun-ignored: reference sources
# never enter this repository's history), applies the five compile-leniency
# patches SURVEY.md section 8c lists (MSVC accepts them, GCC 13 does not; none changes
# arithmetic), compiles every translation unit on the render path with nvcc
# for sm_100a and links them with OUR headless driver (oracle/ref_driver.cu)
# and OUR 11-symbol SDL stub (oracle/sdl_stub.cpp) into oracle/_ref/ref_render.
#
# The binary can only RUN on a GPU box (it is the reference's CUDA renderer).
# It is used by tests/ (-m gpu) and tools/ to pin the CPU oracle and the
# product against the real thing.  The product never links or executes it.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${REF:-/root/reference/CustomCUDARasterizer}"
SDLINC="${SDLINC:-/root/reference/3rdParty/include}"
OUT="$HERE/_ref"
SRC="$OUT/src"
OBJ="$OUT/obj"
NVCC="${NVCC:-nvcc}"

if [ ! -d "$REF" ]; then
  echo "build_ref: $REF not present (GPU box?) - keeping prebuilt $OUT/ref_render" >&2
  exit 0
fi

mkdir -p "$SRC" "$OBJ" "$OUT/resources"
# the bundled meshes and textures of BASELINE configs 1/2/5 (data, not sources): git-ignored with the rest of
# oracle/_ref, but they travel to the GPU box so the parity tests can render the reference's own scenes
cp "$REF"/Resources/*.obj "$REF"/Resources/*.png "$OUT/resources"/
cp "$REF"/*.h "$REF"/*.cpp "$REF"/*.cu "$REF"/*.cuh "$SRC"/

# -- patch 1: GCC's <cmath> has no std::powf (RGBColor.h:126-128)
sed -i 's/std::powf/powf/g' "$SRC/RGBColor.h"
# -- patch 2+3: flexible array members inside unions / non-trivial members in an
#    anonymous struct (Structs.h:6-18 RenderData is unused -> dropped; :45 sized)
python3 - "$SRC/Structs.h" <<'EOF'
import re, sys
p = sys.argv[1]
s = open(p, encoding="latin-1").read()
s = re.sub(r"struct RenderData.*?\n};\n", "", s, count=1, flags=re.S)
s = s.replace("unsigned int idx[];", "unsigned int idx[3];")
open(p, "w", encoding="latin-1").write(s)
EOF
# -- patch 4: member-name typo in the unused single-queue class (CUDAAtomicQueue.cuh:105,112)
sed -i 's/m_Dev_QueueSizeMutex\b/m_Dev_QueueSizeMutexBuffer/g' "$SRC/CUDAAtomicQueue.cuh"
# -- patch 5: reinterpret_cast of an rvalue (CUDAROPs.cu:79-83) -> named temporary
python3 - "$SRC/CUDAROPs.cu" <<'EOF'
import sys
p = sys.argv[1]
s = open(p, encoding="latin-1").read()
old = """	return reinterpret_cast<RGBColor&>(InterpolateVector(
		reinterpret_cast<const FVector3&>(v0),
		reinterpret_cast<const FVector3&>(v1),
		reinterpret_cast<const FVector3&>(v2),
		weights, invDepths));"""
new = """	FVector3 tmp = InterpolateVector(
		reinterpret_cast<const FVector3&>(v0),
		reinterpret_cast<const FVector3&>(v1),
		reinterpret_cast<const FVector3&>(v2),
		weights, invDepths);
	return reinterpret_cast<RGBColor&>(tmp);"""
assert old in s or old.replace("\n", "\r\n") in s, "patch 5 site not found"
s = s.replace(old.replace("\n", "\r\n"), new.replace("\n", "\r\n")).replace(old, new)
open(p, "w", encoding="latin-1").write(s)
EOF

FLAGS=(-std=c++17 -O2 -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr -lineinfo
       -I"$SRC" -I"$SDLINC/sdl2-2.0.9" -I"$SDLINC/sdl2_image-2.0.5" -w)

# the reference's render-path translation units (main.cpp/ERenderer.cpp/ETimer.cpp are not on it)
"$NVCC" "${FLAGS[@]}" -c "$SRC/CUDARenderer.cu" -o "$OBJ/CUDARenderer.o" &
for f in Camera CUDAMesh CUDATexture CUDATextureManager Mesh ObjParser SceneGraph SceneManager \
         Resolution CUDABenchMarker Light DirectionalLight PointLight EventManager; do
  "$NVCC" "${FLAGS[@]}" -x cu -c "$SRC/$f.cpp" -o "$OBJ/$f.o" &
done
wait

# our driver + SDL stub (these are NOT reference code)
"$NVCC" "${FLAGS[@]}" -x cu -c "$HERE/ref_driver.cu" -o "$OBJ/ref_driver.o"
g++ -O2 -I"$SDLINC/sdl2-2.0.9" -I"$SDLINC/sdl2_image-2.0.5" -c "$HERE/sdl_stub.cpp" -o "$OBJ/sdl_stub.o"

"$NVCC" -gencode arch=compute_100a,code=sm_100a "$OBJ"/*.o -o "$OUT/ref_render"
echo "build_ref: built $OUT/ref_render"
