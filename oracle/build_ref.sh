#!/usr/bin/env bash
# oracle/build_ref.sh -- TEST INFRASTRUCTURE ONLY.
# Compiles the reference's OWN C sources, unmodified, from where they lie under /root/reference
# (never copied into this repo) against the shim headers in oracle/shims/ (FFTW3, MPI and GSL are
# absent from this image).  Outputs go only to oracle/_ref/ (git-ignored, but shipped to the GPU box):
#   oracle/_ref/boltz_      the reference driver (exec/boltz.c), 1 MPI rank, OpenMP
#   oracle/_ref/libref.so   the reference's hot-path functions, for function-level parity via ctypes
# Flags follow the reference's Release default (CMakeLists.txt:6-10,22-39,62-65) with gnu99 instead
# of c99 (weights.c uses M_PI without constants.h).  Never -O0: the OpenMP loop at
# src/collisions.c:127 races on xi/index unless they live in registers.
set -euo pipefail
REF="${SBTE_REFERENCE_ROOT:-/root/reference}"
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
if [ ! -d "$REF/src" ]; then
  echo "build_ref.sh: $REF not present; keeping whatever is already in $OUT" >&2
  exit 0
fi
mkdir -p "$OUT"
CFLAGS="-std=gnu99 -O3 -fopenmp -fPIC -w -I$HERE/shims -I$REF/src -I$REF/exec"
SRC_COMMON="$REF/src/collisions.c $REF/src/conserve.c $REF/src/momentRoutines.c \
  $REF/src/transportroutines.c $REF/src/boundaryConditions.c $REF/src/weights.c"
SRC_DRIVER="$REF/exec/boltz.c $REF/src/initializer.c $REF/src/input.c $REF/src/output.c \
  $REF/src/mesh_setup.c $REF/src/restart.c $REF/src/species.c"
gcc $CFLAGS -o "$OUT/boltz_" $SRC_DRIVER $SRC_COMMON "$HERE/shims/shim.c" "$HERE/qag21.c" -lm
gcc $CFLAGS -shared -o "$OUT/libref.so" $SRC_COMMON $REF/src/initializer.c $REF/src/restart.c \
  $REF/src/species.c $REF/src/input.c "$HERE/shims/shim.c" "$HERE/qag21.c" -lm
# The reference driver linked against the B200 library INSTEAD of its own collision / conservation /
# transport modules (the drop-in claim of include/sbte_b200.h, exercised by tests/test_gpu_dropin_driver.py)
LIBDIR="$HERE/../spectralbte_b200"
if [ -f "$LIBDIR/libsbte_b200.so" ]; then
  gcc $CFLAGS -o "$OUT/boltz_gpu" $SRC_DRIVER $REF/src/momentRoutines.c $REF/src/weights.c \
    "$HERE/shims/shim.c" "$HERE/qag21.c" -L"$LIBDIR" -lsbte_b200 -Wl,-rpath,'$ORIGIN/../../spectralbte_b200' -lm
fi
echo "built $OUT/boltz_ and $OUT/libref.so"
