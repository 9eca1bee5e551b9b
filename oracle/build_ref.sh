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
# Drop-in variant: a config written as "<name>+b200" additionally replaces src/rhs.cu by
# integration/rhs_b200.cu and links libb200sph_<name>.so, i.e. the reference host (main, I/O,
# libconfig, integrators) running on the new kernels -> oracle/_ref/miluphcuda_<name>_b200.
#
# usage: oracle/build_ref.sh [config[+b200] ...]     (default: all six, reference only)
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
  [nakamura]="test_cases/nakamura"
)

if [ ! -d "$REF/src" ]; then
  echo "build_ref: $REF not present (expected on the GPU box); using prebuilt oracle/_ref" >&2
  exit 0
fi

CONFIGS=("$@")
if [ ${#CONFIGS[@]} -eq 0 ]; then CONFIGS=(shocktube sedov rings impact giant_hydro giant_solid nakamura); fi

mkdir -p "$OUT"
# The reference's SHIPPED inputs (the parity fixtures the reference itself provides) and its shipped
# parameter.h files are staged next to the binaries: git-ignored like them, they travel to the GPU box with the
# snapshot and are read there by tests/test_gpu_live_reference.py and tests/test_shipped_parameter_h.py.
stage_fixture() {  # <name> <reference dir> <files...>
  local name="$1" dir="$2"; shift 2
  mkdir -p "$OUT/fixtures/$name"
  for f in "$@"; do
    if [ -f "$REF/$dir/$f" ] && [ ! -f "$OUT/fixtures/$name/$f" ]; then install -m 0644 "$REF/$dir/$f" "$OUT/fixtures/$name/$f"; fi
  done
}
stage_fixture impact examples/impact impact.0000.gz material.cfg parameter.h
stage_fixture giant_hydro examples/giant_collisions/hydro impact.0000.gz material.cfg iron.till.cfg granite.till.cfg parameter.h
stage_fixture giant_solid examples/giant_collisions/solid impact.0000.gz material.cfg iron.till.cfg granite.till.cfg parameter.h
stage_fixture shocktube test_cases/shocktube parameter.h material.cfg
stage_fixture sedov test_cases/sedov parameter.h material.cfg
stage_fixture rings test_cases/colliding_rings parameter.h material.cfg
stage_fixture nakamura test_cases/nakamura parameter.h material.cfg
for spec in "${CONFIGS[@]}"; do
  cfg="${spec%+b200}"
  dropin=0; [ "$spec" != "$cfg" ] && dropin=1
  suffix=""; [ $dropin -eq 1 ] && suffix="_b200"
  src_dir="${CFGDIR[$cfg]:-}"
  if [ -z "$src_dir" ]; then echo "unknown config $cfg" >&2; exit 2; fi
  B="$OUT/build/$cfg$suffix"
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

  NVFLAGS="$ARCH -x cu -c -dc -O3 -w -Xcompiler -O3,-pthread -DVERSION=\"ref-sm100a\" -I$B/include -I$REPO/miluphcuda_b200/csrc -I$REPO/include"
  echo "[build_ref] $cfg: compiling"
  ( for f in "$REF"/src/*.cu; do
      n="$(basename "$f" .cu)"
      [ "$n" = "euler" ] && continue
      if [ $dropin -eq 1 ] && [ "$n" = "rhs" ]; then continue; fi
      echo "$f $B/obj/$n.o"
    done
    echo "$HERE/ref_hook.cu $B/obj/ref_hook.o"
    if [ $dropin -eq 1 ]; then echo "$REPO/integration/rhs_b200.cu $B/obj/rhs_b200.o"; fi
  ) | xargs -P "$(nproc)" -n 2 sh -c "$NVCC $NVFLAGS -o \"\$1\" \"\$0\""
  gcc -O2 -c "$REPO/miluphcuda_b200/csrc/libconfig_lite.c" -o "$B/obj/libconfig_lite.o"
  LINK_EXTRA=""
  if [ $dropin -eq 1 ]; then
    python3 "$REPO/miluphcuda_b200/build.py" "$cfg" > /dev/null
    LINK_EXTRA="-L$REPO/miluphcuda_b200/lib -lb200sph_$cfg -Xlinker -rpath -Xlinker \$ORIGIN/../../miluphcuda_b200/lib"
  fi
  $NVCC $ARCH "$B"/obj/*.o -lcudart -lpthread $LINK_EXTRA -o "$OUT/miluphcuda_$cfg$suffix"
  rm -rf "$B"
  echo "[build_ref] $spec: $OUT/miluphcuda_$cfg$suffix"
done
rmdir "$OUT/build" 2>/dev/null || true
