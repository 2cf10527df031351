#!/bin/bash
# Builds libgpet_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
OUT="$HERE/../libgpet_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -I"$ROOT/include" -I"$HERE")
OBJ="$(mktemp -d)"
trap 'rm -rf "$OBJ"' EXIT
# the L-BFGS-B state machines must not be contracted into fused multiply-adds (see gpet_lbfgsb.cuh)
"$NVCC" "${FLAGS[@]}" -fmad=false -Xcompiler -ffp-contract=off ${GPET_NVCC_EXTRA:-} -c "$HERE"/gpet_lbfgsb.cu -o "$OBJ"/gpet_lbfgsb.o
"$NVCC" "${FLAGS[@]}" -shared ${GPET_NVCC_EXTRA:-} \
    "$HERE"/gpet_cabi.cu "$HERE"/gpet_image.cu "$HERE"/gpet_posterior.cu "$HERE"/gpet_factor.cu \
    "$HERE"/gpet_sample.cu "$HERE"/gpet_score.cu "$HERE"/gpet_density.cu "$HERE"/gpet_finalfit.cu \
    "$HERE"/gpet_rng.cu "$OBJ"/gpet_lbfgsb.o \
    -o "$OUT"
echo "built $OUT"
