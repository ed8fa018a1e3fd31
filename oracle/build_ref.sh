#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY -- builds the UNMODIFIED reference (felpzOliveira/Bubbles) CPU path
# from the sources where they lie under $BUBBLES_REF (default /root/reference) into oracle/_ref/.
# Nothing is copied into the repo; only objects + the harness binary land in oracle/_ref/ (git-ignored).
# Recipe follows SURVEY.md Appendix D (the reference's own CMake build needs a GPU probe, X11 and Qhull).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${BUBBLES_REF:-/root/reference}"
OUT="$HERE/_ref"
OBJ="$OUT/obj"
JOBS="${JOBS:-8}"
if [ ! -d "$REF/src" ]; then
  echo "[build_ref] reference sources not present at $REF; keeping prebuilt $OUT (if any)"; exit 0
fi
mkdir -p "$OBJ"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
INC=""
for d in core shapes third third/graphy cuda tests apps reconstruction boundaries; do INC="$INC -I$REF/src/$d"; done
FLAGS="-x cu -dc -std=c++17 -O2 -DRELEASE --use_fast_math --extended-lambda --gpu-architecture=sm_100 -w $INC"
SRCS=$(ls $REF/src/core/*.cpp $REF/src/cuda/cutil.cpp \
  $REF/src/equations/sph_equations2.cpp $REF/src/equations/sph_equations3.cpp $REF/src/equations/pcisph_equations3.cpp \
  $REF/src/generator/*.cpp $REF/src/shapes/*.cpp \
  $REF/src/solvers/sph_solver3.cpp $REF/src/solvers/pcisph_solver3.cpp \
  $REF/src/reconstruction/*.cpp $REF/src/third/*.cpp)
compile_one() {
  src="$1"; name=$(echo "$src" | sed "s#$REF/src/##; s#/#_#g; s#\.cpp\$#.o#")
  if [ ! -f "$OBJ/$name" ] || [ "$src" -nt "$OBJ/$name" ]; then
    $NVCC $FLAGS -c "$src" -o "$OBJ/$name" || { echo "[build_ref] FAILED $src"; exit 1; }
  fi
}
export -f compile_one; export REF OBJ NVCC FLAGS
echo "$SRCS" | xargs -P "$JOBS" -I{} bash -c 'compile_one {}'
# compile check of the reference-side binding against the unmodified reference headers + include/bbx.h
BIND="$HERE/../bubbles_b200/host/reference_binding/pcisph_solver3_bbx.cpp"
if [ -f "$BIND" ]; then
  mkdir -p "$OUT/obj_bbx"
  if [ ! -f "$OUT/obj_bbx/pcisph_solver3_bbx.o" ] || [ "$BIND" -nt "$OUT/obj_bbx/pcisph_solver3_bbx.o" ] || [ "$HERE/../include/bbx.h" -nt "$OUT/obj_bbx/pcisph_solver3_bbx.o" ]; then
    $NVCC $FLAGS -I"$HERE/../include" -c "$BIND" -o "$OUT/obj_bbx/pcisph_solver3_bbx.o" || { echo "[build_ref] FAILED reference binding"; exit 1; }
    echo "[build_ref] reference-side binding compiles against the reference headers"
  fi
fi
# nothing to do when both binaries are newer than the harness source and every reference object
NEWEST_OBJ=$(ls -t "$OBJ"/*.o 2>/dev/null | grep -v "/_harness.o" | head -1 || true)
if [ -x "$OUT/bbref" ] && [ -x "$OUT/bbref_gpu" ] && [ "$OUT/bbref" -nt "$HERE/ref_harness.cpp" ] && [ "$OUT/bbref_gpu" -nt "$HERE/ref_harness.cpp" ] \
   && [ -n "$NEWEST_OBJ" ] && [ "$OUT/bbref" -nt "$NEWEST_OBJ" ] && [ "$OUT/bbref_gpu" -nt "$NEWEST_OBJ" ] && [ "$OUT/bbref" -nt "$HERE/build_ref.sh" ]; then
  echo "[build_ref] up to date: $OUT/bbref, $OUT/bbref_gpu"; exit 0
fi
# harness (ours) + malloc shim for the managed-memory arena
$NVCC $FLAGS -c "$HERE/ref_harness.cpp" -o "$OBJ/_harness.o"
$NVCC --gpu-architecture=sm_100 -o "$OUT/bbref" "$OBJ"/*.o -ldl -lpthread
echo "[build_ref] built $OUT/bbref"
# the reference's GPU path (its native mode) with the same driver: its own managed-memory arena instead of the shim
mkdir -p "$OUT/obj_gpu"
$NVCC $FLAGS -c "$REF/src/cuda/memory.cpp" -o "$OUT/obj_gpu/cuda_memory.o"
$NVCC $FLAGS -DBBREF_GPU -c "$HERE/ref_harness.cpp" -o "$OUT/obj_gpu/_harness_gpu.o"
$NVCC --gpu-architecture=sm_100 -o "$OUT/bbref_gpu" $(ls "$OBJ"/*.o | grep -v "/_harness.o") "$OUT/obj_gpu/cuda_memory.o" "$OUT/obj_gpu/_harness_gpu.o" -ldl -lpthread
echo "[build_ref] built $OUT/bbref_gpu"
