#!/usr/bin/env bash
# TEST INFRASTRUCTURE -- builds the UNMODIFIED reference (miluphcuda) for sm_100a
# straight from /root/reference into oracle/_ref/ (git-ignored, travels to the
# GPU box).  No reference source is copied into the repository: sources are
# compiled where they lie; the per-config include directory is a farm of
# symlinks plus two generated headers that live only in the scratch build dir
# and are deleted after linking.
#
# Shims:
#   * fast_integration.h  -- empty stub (file missing from the snapshot)
#   * libconfig.h         -- forwards to miluphcuda_b200/csrc/libconfig_lite.h
#   * parameter.h         -- the scenario's own file with HDF5IO/MORE_OUTPUT set to 0
#   * miluph.h            -- "quiet" variant: DEBUG_TIMESTEP/TREE/GRAVITY/RHS 0,
#                            DEBUG_RHS_RUNTIMES stays 1 (per-kernel durations)
#   * src/euler.cu is replaced by oracle/ref_hook.cu (dump + timing hook that
#     calls the reference's own rightHandSide())
#
# usage: oracle/build_ref.sh [config ...]     (default: all six)
set -euo pipefail

HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REPO="$(dirname "$HERE")"
REF="${REFERENCE_ROOT:-/root/reference}"
OUT="$HERE/_ref"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
ARCH="-gencode arch=compute_100a,code=sm_100a"

declare -A CFGDIR=(
  [shocktube]="test_cases/shocktube"
  [sedov]="test_cases/sedov"
  [rings]="test_cases/colliding_rings"
  [impact]="examples/impact"
  [giant_hydro]="examples/giant_collisions/hydro"
  [giant_solid]="examples/giant_collisions/solid"
)

if [ ! -d "$REF/src" ]; then
  echo "build_ref: $REF not present (expected on the GPU box); using prebuilt oracle/_ref" >&2
  exit 0
fi

CONFIGS=("$@")
if [ ${#CONFIGS[@]} -eq 0 ]; then CONFIGS=(shocktube sedov rings impact giant_hydro giant_solid); fi

mkdir -p "$OUT"
for cfg in "${CONFIGS[@]}"; do
  src_dir="${CFGDIR[$cfg]:-}"
  if [ -z "$src_dir" ]; then echo "unknown config $cfg" >&2; exit 2; fi
  B="$OUT/build/$cfg"
  rm -rf "$B"; mkdir -p "$B/include" "$B/obj"
  for h in "$REF"/include/*.h; do
    n="$(basename "$h")"
    case "$n" in parameter.h|miluph.h) ;; *) ln -s "$h" "$B/include/$n";; esac
  done
  ln -s "$HERE/shim/fast_integration.h" "$B/include/fast_integration.h"
  ln -s "$HERE/shim/libconfig.h" "$B/include/libconfig.h"
  sed -E 's/^(#define[[:space:]]+HDF5IO)[[:space:]]+1/\1 0/; s/^(#define[[:space:]]+MORE_OUTPUT)[[:space:]]+1/\1 0/' \
      "$REF/$src_dir/parameter.h" > "$B/include/parameter.h"
  sed -E 's/^(#define[[:space:]]+DEBUG_(TIMESTEP|TREE|GRAVITY|RHS))[[:space:]]+1/\1 0/' \
      "$REF/include/miluph.h" > "$B/include/miluph.h"

  NVFLAGS="$ARCH -x cu -c -dc -O3 -w -Xcompiler -O3,-pthread -DVERSION=\"ref-sm100a\" -I$B/include -I$REPO/miluphcuda_b200/csrc"
  echo "[build_ref] $cfg: compiling"
  ( for f in "$REF"/src/*.cu; do
      n="$(basename "$f" .cu)"
      [ "$n" = "euler" ] && continue
      echo "$f $B/obj/$n.o"
    done
    echo "$HERE/ref_hook.cu $B/obj/ref_hook.o"
  ) | xargs -P "$(nproc)" -n 2 sh -c "$NVCC $NVFLAGS -o \"\$1\" \"\$0\""
  gcc -O2 -c "$REPO/miluphcuda_b200/csrc/libconfig_lite.c" -o "$B/obj/libconfig_lite.o"
  $NVCC $ARCH "$B"/obj/*.o -lcudart -lpthread -o "$OUT/miluphcuda_$cfg"
  rm -rf "$B"
  echo "[build_ref] $cfg: $OUT/miluphcuda_$cfg"
done
rmdir "$OUT/build" 2>/dev/null || true
