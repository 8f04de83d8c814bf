#!/usr/bin/env bash
# Builds the libstage binding and its test driver against the Stage headers where they lie under /root/reference:
#   integration/_build/libstage_gpu.so   the interposing binding (integration/stage_gpu_shim.cc), links libstg_rt.so
#   integration/_build/stage_run         headless driver that dumps what controllers can read (integration/stage_run.cc)
# Needs oracle/_ref (the unmodified reference built headless by oracle/build_ref.sh) to link and run against.
# Outputs only into integration/_build/ (git-ignored; travels to the GPU box with gpurun). Where /root/reference is
# absent (the GPU box) this keeps the prebuilt files.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(dirname "$HERE")"
REF="${STG_REFERENCE_DIR:-/root/reference}"
OUT="$HERE/_build"
if [ ! -d "$REF/libstage" ]; then
  echo "integration/build.sh: $REF not present; keeping prebuilt $OUT" >&2
  exit 0
fi
[ -f "$ROOT/oracle/_ref/libstage_ref.so" ] || "$ROOT/oracle/build_ref.sh"
mkdir -p "$OUT"
CXX="${CXX:-g++}"
FLAGS="-std=gnu++11 -O2 -DNDEBUG -fPIC -w -I$ROOT/oracle/ref_shim -I$REF/libstage -I$REF/replace -I$ROOT/include"
# the binding: resolves libstg_rt.so next to the package (rpath), libstage at run time (it is preloaded before it)
$CXX $FLAGS -fno-access-control -shared "$HERE/stage_gpu_shim.cc" -o "$OUT/libstage_gpu.so" \
  -L"$ROOT/stage_b200" -lstg_rt -Wl,-rpath,'$ORIGIN/../../stage_b200' -ldl -lpthread
# the driver: -rdynamic so that its World::GetEventQueue is found before libstage's
$CXX $FLAGS -rdynamic "$HERE/stage_run.cc" -o "$OUT/stage_run" -L"$ROOT/oracle/_ref" -lstage_ref \
  -Wl,-rpath,'$ORIGIN/../../oracle/_ref' -ldl -lpthread
echo "integration/build.sh: ok -> $OUT"
