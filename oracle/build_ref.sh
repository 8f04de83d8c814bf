#!/usr/bin/env bash
# Build the UNMODIFIED reference (rtv/Stage 4.3.0) headless from the sources where they lie
# under /root/reference, against the stub FLTK/GL/ltdl headers in oracle/ref_shim/.
# Outputs go ONLY to oracle/_ref/ (git-ignored; travels to the GPU box with gpurun).
# This is test infrastructure: nothing under stage_b200/ links or loads it.
# The reference's own build system (cmake + FLTK + GL + ltdl) is not run; see DESIGN.md.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${STG_REFERENCE_DIR:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$REF/libstage" ]; then
  echo "build_ref: $REF not present; keeping prebuilt $OUT" >&2
  exit 0
fi
mkdir -p "$OUT/obj" "$OUT/pgm" "$OUT/plugins" "$OUT/share/stage" "$OUT/worlds"
CXX="${CXX:-g++}"
# same optimisation level as the reference's release build (CMakeLists.txt:53-55): -O2, no -march
FLAGS="-std=gnu++11 -O2 -DNDEBUG -fPIC -w -I$HERE/ref_shim -I$REF/libstage -I$REF/replace"
SRCS="world region block blockgroup model model_ranger model_fiducial model_blobfinder
      model_position ancestor model_callbacks worldfile color powerpack logentry file_manager
      model_bumper model_gripper model_actuator model_blinkenlight model_lightindicator stage"
pids=()
for s in $SRCS; do
  if [ ! -f "$OUT/obj/$s.o" ] || [ "$REF/libstage/$s.cc" -nt "$OUT/obj/$s.o" ]; then
    $CXX $FLAGS -c "$REF/libstage/$s.cc" -o "$OUT/obj/$s.o" &
    pids+=($!)
  fi
done
$CXX $FLAGS -c "$HERE/ref_stubs.cc" -o "$OUT/obj/ref_stubs.o" &
pids+=($!)
for p in "${pids[@]}"; do wait "$p"; done
OBJS=""
for s in $SRCS ref_stubs; do OBJS="$OBJS $OUT/obj/$s.o"; done
$CXX -shared -o "$OUT/libstage_ref.so" $OBJS -lpthread -ldl
# the reference's own front-end and the controllers its shipped worlds name
$CXX $FLAGS "$REF/libstage/main.cc" -o "$OUT/stage" -L"$OUT" -lstage_ref -Wl,-rpath,'$ORIGIN' -lpthread -ldl
for c in wander fasr source sink pioneer_flocking lasernoise; do
  $CXX $FLAGS -shared "$REF/examples/ctrl/$c.cc" -o "$OUT/plugins/$c.so" -L"$OUT" -lstage_ref -Wl,-rpath,'$ORIGIN/..'
done
cp "$REF/libstage/rgb.txt" "$OUT/share/stage/rgb.txt"
# world files are inputs (config data), staged so the harness also runs where /root/reference is absent
cp -r "$REF/worlds/." "$OUT/worlds/"
cp "$REF/libstageplugin/test/lsp_test.world" "$OUT/worlds/lsp_test.world" 2>/dev/null || true
python3 - "$OUT/worlds/bitmaps" <<'PY'
import sys, glob, os
from PIL import Image
for p in glob.glob(os.path.join(sys.argv[1], "*.png")):
    im = Image.open(p).convert("L")
    with open(p + ".pgm", "wb") as f:
        f.write(b"P5 %d %d 255\n" % im.size)
        f.write(im.tobytes())
PY
# the interposing harness (oracle/ref_harness.cc): a shared library python loads via ctypes
# -fno-access-control: the harness reads protected members of the reference's classes (Model::pose, MapWithChildren)
$CXX $FLAGS -fno-access-control -shared "$HERE/ref_harness.cc" -o "$OUT/libref_harness.so" -L"$OUT" -lstage_ref -Wl,-rpath,'$ORIGIN' -lpthread -ldl
echo "build_ref: ok -> $OUT"
