#!/usr/bin/env bash
# TEST INFRASTRUCTURE -- builds oracle/_ref/dropin_main: the reference's own scene-construction code (AddTexture and
# CreateScenes of CustomCUDARasterizer/main.cpp, extracted verbatim at build time into a git-ignored scratch file; its
# commented-out scene blocks 1, 2, 3 and 5 are switched on by removing the leading `//`) compiled against the PRODUCT's
# C++ veneer and linked with libb2r.so.  It shows that a maintainer's existing host code compiles and runs unchanged on
# the new renderer (tests/test_cpu_host.py: the scene graphs it builds; tests/test_gpu_veneer.py: the frames).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${REF:-/root/reference/CustomCUDARasterizer}"
CSRC="$HERE/../custom_gpgpu_rasterizer_b200/csrc"
OUT="$HERE/_ref"
SRC="$OUT/src_dropin"

if [ ! -f "$REF/main.cpp" ]; then
  echo "build_ref_dropin: $REF not present (GPU box?) - keeping prebuilt $OUT/dropin_main" >&2
  exit 0
fi
mkdir -p "$SRC"
python3 - "$REF/main.cpp" "$SRC/create_scenes.inc" <<'PY'
import sys
lines = open(sys.argv[1], encoding="utf-8-sig", errors="replace").read().replace("\r\n", "\n").split("\n")
a = next(i for i, l in enumerate(lines) if l.startswith("int AddTexture"))
b = next(i for i, l in enumerate(lines) if l.startswith("void ShutDown"))
body = lines[a:b]
for title in ("// SceneGraph 1", "// SceneGraph 2", "SceneGraph 3 // TukTuk", "SceneGraph 5 // Vehicle"):
    i = next(k for k, l in enumerate(body) if l.rstrip().endswith(title))
    s = i - 1
    assert body[s].strip() == "//{", body[s]
    e = next(k for k in range(i, len(body)) if body[k].rstrip() == "\t//}")
    for k in range(s, e + 1):
        body[k] = body[k].replace("//", "", 1)
open(sys.argv[2], "w").write("\n".join(body) + "\n")
PY
"${CXX:-g++}" -std=c++17 -O2 -ffp-contract=off -w "$HERE/dropin_main.cpp" -I"$SRC" -I"$CSRC/host" -L"$CSRC" -lb2r \
  -Wl,-rpath,"$(cd "$CSRC" && pwd)" -o "$OUT/dropin_main"
echo "build_ref_dropin: built $OUT/dropin_main"
