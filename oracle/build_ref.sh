#!/usr/bin/env bash
# oracle/build_ref.sh -- TEST INFRASTRUCTURE, NOT PRODUCT.
#
# Compiles the UNMODIFIED reference solver sources where they lie under
# /root/reference/FV (never copied into this repo) against oracle/shim/mpi.h,
# with outputs only into oracle/_ref/ (git-ignored, travels to the GPU box):
#   _ref/MPI2DFV          run-time configurable binary (ORACLE_NITER/IAXI/IGAM/ACCUR)
#   _ref/MPI2DFV_strict   header differs from the reference's only in NITER
#   _ref/libpar2d_ref.so  reference functions behind extern "C" (ref_wrap.cpp)
#   _ref/GridPart         the reference's grid partitioner, unmodified (GridPart/main.c: its i split is the
#                         hard-coded node table istepmanual[] = {0, 99, 178, 257, 356})
#   _ref/GridPart_auto    the same source with the automatic i split it keeps in comments restored by sed
#                         (generated copy under _ref/gen_gridpart/, a build artefact like the headers above)
# Usage: oracle/build_ref.sh [build|verify]
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${PAR2D_REFERENCE:-/root/reference}/FV"
OUT="$HERE/_ref"
CXXFLAGS="-std=c++17 -O2 -w"

if [ ! -d "$REF" ]; then
    echo "build_ref: $REF not present (GPU box?) -- keeping prebuilt $OUT" >&2
    exit 0
fi
mkdir -p "$OUT/gen_rt" "$OUT/gen_strict"

# Generated headers (build artefacts, not committed): the reference header with
# three #define lines redirected.  g++ -include makes them win over the copy
# next to the sources because they claim the same include guard first.
{
    echo '#include "ref_runtime_config.h"'
    sed -e 's|^#define IGAM .*|#define IGAM (oracle_ref_config().igam)|' \
        -e 's|^#define IAXI .*|#define IAXI (oracle_ref_config().iaxi)|' \
        -e 's|^#define ACCUR .*|#define ACCUR (oracle_ref_config().accur)|' \
        -e 's|^#define NITER .*|#define NITER (oracle_ref_config().niter)|' "$REF/Indexing.h"
} > "$OUT/gen_rt/Indexing.h"
sed -e 's|^#define NITER .*|#define NITER (25)|' "$REF/Indexing.h" > "$OUT/gen_strict/Indexing.h"

SOLVER_TUS="EulerFlux BoundaryConditions FileIO SpatialDiscretization MeshModule LUtools Jacobian Thermo DGP1Tools"
srcs=""
for t in $SOLVER_TUS; do srcs="$srcs $REF/$t.cpp"; done

build() {
    g++ $CXXFLAGS -I"$HERE" -I"$HERE/shim" -include "$OUT/gen_rt/Indexing.h" -I"$REF" \
        "$REF/main.cpp" $srcs -o "$OUT/MPI2DFV"
    g++ $CXXFLAGS -I"$HERE/shim" -include "$OUT/gen_strict/Indexing.h" -I"$REF" \
        "$REF/main.cpp" $srcs -o "$OUT/MPI2DFV_strict"
    g++ $CXXFLAGS -fPIC -shared -I"$HERE" -I"$HERE/shim" -include "$OUT/gen_rt/Indexing.h" -I"$REF" \
        "$HERE/ref_wrap.cpp" $srcs -o "$OUT/libpar2d_ref.so"
    GP="$(dirname "$REF")/GridPart/main.c"
    if [ -f "$GP" ]; then
        mkdir -p "$OUT/gen_gridpart"
        gcc -O2 -w "$GP" -o "$OUT/GridPart" -lm
        sed -e 's|istep = -999;|istep = ceil((double)ni[b]/ipart[b] - 1e-8);|' \
            -e 's|int istart = istepmanual\[iblock\];//istep\*iblock;|int istart = istep*iblock;|' \
            -e 's|int iend = istepmanual\[iblock+1\];//MIN(ni\[b\]-1, istep\*(iblock+1));|int iend = MIN(ni[b]-1, istep*(iblock+1));|' \
            "$GP" > "$OUT/gen_gridpart/main.c"
        grep -q 'int istart = istep\*iblock;' "$OUT/gen_gridpart/main.c" && grep -q 'int iend = MIN(ni\[b\]-1, istep\*(iblock+1));' "$OUT/gen_gridpart/main.c"
        gcc -O2 -w "$OUT/gen_gridpart/main.c" -o "$OUT/GridPart_auto" -lm
    fi
    echo "build_ref: built $(ls "$OUT" | tr '\n' ' ')"
}

# verify: (1) run-time-config binary == strict binary bitwise after 25 steps on
# FV/Case with 8 shim ranks; (2) both reproduce the reference's committed golden
# /root/reference/FV/Outputs/res.tec for the lines produced.
verify() {
    tmp="$(mktemp -d)"
    for v in rt strict; do
        mkdir -p "$tmp/$v/build" "$tmp/$v/Outputs"
        cp -r "$REF/Case" "$tmp/$v/Case"
        cp "$REF/therm_air.dat" "$tmp/$v/"
    done
    (cd "$tmp/rt/build" && SHIM_NP=8 ORACLE_NITER=25 "$OUT/MPI2DFV" > log.txt)
    (cd "$tmp/strict/build" && SHIM_NP=8 "$OUT/MPI2DFV_strict" > log.txt)
    for b in 1 2 3 4 5 6 7 8; do
        cmp "$tmp/rt/Outputs/solution.b$b.tec" "$tmp/strict/Outputs/solution.b$b.tec"
    done
    cmp "$tmp/rt/Outputs/res.tec" "$tmp/strict/Outputs/res.tec"
    n=$(wc -l < "$tmp/rt/Outputs/res.tec")
    head -n "$n" "$REF/Outputs/res.tec" | cmp - "$tmp/rt/Outputs/res.tec"
    echo "build_ref verify: OK ($((n-1)) residual lines identical to the reference's committed res.tec; rt == strict bitwise)"
    rm -rf "$tmp"
}

case "${1:-build}" in
    build) build ;;
    verify) build; verify ;;
    *) echo "usage: $0 [build|verify]" >&2; exit 2 ;;
esac
