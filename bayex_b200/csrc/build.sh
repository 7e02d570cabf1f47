#!/usr/bin/env bash
# Build libbayex_b200.so for sm_100a (in-tree; the .so travels to the GPU box with the snapshot).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../libbayex_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall
       --expt-relaxed-constexpr -Xptxas -v)
UNITS="linalg fit score score_tc16p selftest topk refine"
mkdir -p "$HERE/build"
rm -f "$HERE"/build/*.o
pids=()
for f in $UNITS; do
  ( "$NVCC" "${FLAGS[@]}" -c "$HERE/$f.cu" -o "$HERE/build/$f.o" > "$HERE/build/$f.log" 2>&1 || { cat "$HERE/build/$f.log"; exit 1; } ) &
  pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
OBJS=(); for f in $UNITS; do OBJS+=("$HERE/build/$f.o"); done
"$NVCC" -shared -o "$OUT" "${OBJS[@]}"
echo "built $OUT"
