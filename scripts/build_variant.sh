#!/usr/bin/env bash
# Dev tool: build a tuning variant of the library (extra -D flags for one translation unit) into csrc/build/variants/<name>.so
# usage: scripts/build_variant.sh <name> <unit> "<-D flags>"     then run with BAYEX_B200_LIB=<path>
set -euo pipefail
NAME="$1"; UNIT="$2"; DEFS="$3"
CS="$(cd "$(dirname "${BASH_SOURCE[0]}")/../bayex_b200/csrc" && pwd)"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
mkdir -p "$CS/build/variants"
"$NVCC" -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr $DEFS \
  -c "$CS/$UNIT.cu" -o "$CS/build/variants/${UNIT}_$NAME.o" 2>/dev/null
OBJS=()
for f in linalg fit score score_tc16p selftest topk refine; do
  if [ "$f" == "$UNIT" ]; then OBJS+=("$CS/build/variants/${UNIT}_$NAME.o"); else OBJS+=("$CS/build/$f.o"); fi
done
"$NVCC" -shared -o "$CS/build/variants/$NAME.so" "${OBJS[@]}"
echo "$CS/build/variants/$NAME.so"
