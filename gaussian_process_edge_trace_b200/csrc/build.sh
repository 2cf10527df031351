#!/bin/bash
# Builds libgpet_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU). Objects compile in parallel; the
# library is linked to a temporary name and renamed into place, so a concurrent loader never sees a partial file.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
OUT="$HERE/../libgpet_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -I"$ROOT/include" -I"$HERE")
OBJ="$(mktemp -d)"
trap 'rm -rf "$OBJ"' EXIT
SRCS=(gpet_cabi gpet_image gpet_posterior gpet_factor gpet_sample gpet_sample_score gpet_score gpet_density gpet_density_bands gpet_finalfit
      gpet_rng gpet_control gpet_testimg gpet_lbfgsb gpet_dense gpet_jacobi)
pids=()
for s in "${SRCS[@]}"; do
    extra=()
    # the L-BFGS-B state machines must not be contracted into fused multiply-adds (see gpet_lbfgsb.cuh)
    if [ "$s" = gpet_lbfgsb ]; then extra=(-fmad=false -Xcompiler -ffp-contract=off); fi
    "$NVCC" "${FLAGS[@]}" "${extra[@]}" ${GPET_NVCC_EXTRA:-} -c "$HERE/$s.cu" -o "$OBJ/$s.o" &
    pids+=($!)
done
fail=0
for p in "${pids[@]}"; do wait "$p" || fail=1; done      # let every compile finish before the temp dir goes away
[ "$fail" = 0 ] || { echo "build.sh: a source file failed to compile" >&2; exit 1; }
objs=()
for s in "${SRCS[@]}"; do objs+=("$OBJ/$s.o"); done
TMP="$OUT.tmp.$$"
"$NVCC" "${FLAGS[@]}" -shared "${objs[@]}" -o "$TMP"
mv -f "$TMP" "$OUT"
echo "built $OUT"
