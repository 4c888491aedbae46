#!/usr/bin/env bash
# Build libpinb200.so (sm_100a) in-tree.  Usage: build.sh [extra nvcc flags]
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
out="$here/../libpinb200.so"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
     -shared -Xcompiler -fPIC -Xptxas -v "$@" \
     -o "$out" "$here"/*.cu 2> "$here/../build.log" || { cat "$here/../build.log"; exit 1; }
grep -E "error|warning: v|spill|Used" "$here/../build.log" | grep -vE "^$" | tail -60 || true
echo "built $out"
