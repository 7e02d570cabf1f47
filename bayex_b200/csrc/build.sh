#!/usr/bin/env bash
# Build libbayex_b200.so for sm_100a (in-tree; the .so travels to the GPU box with the snapshot).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../libbayex_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall
       --expt-relaxed-constexpr -Xptxas -v)
mkdir -p "$HERE/build"
pids=()
for f in linalg fit score score_tc score_tc2 score_tc16 score_tc16p score_tc16s score_tc16t topk refine; do
  ( "$NVCC" "${FLAGS[@]}" -c "$HERE/$f.cu" -o "$HERE/build/$f.o" > "$HERE/build/$f.log" 2>&1 || { cat "$HERE/build/$f.log"; exit 1; } ) &
  pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
"$NVCC" -shared -o "$OUT" "$HERE"/build/{linalg,fit,score,score_tc,score_tc2,score_tc16,score_tc16p,score_tc16s,score_tc16t,topk,refine}.o
echo "built $OUT"
