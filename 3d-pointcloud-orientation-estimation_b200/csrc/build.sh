#!/usr/bin/env bash
# Builds libpcoe.so in-tree for sm_100a.  Usage: csrc/build.sh [extra nvcc flags]
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
root="$(cd "$here/../.." && pwd)"
out="$here/../libpcoe.so"
mkdir -p "$here/build"
NVCC="${NVCC:-nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Wno-deprecated-gpu-targets
       -I"$root/include" -I"$here" "$@")
objs=()
pids=()
# newest header (any change to a shared header rebuilds every object)
newest_hdr="$(ls -t "$here"/*.cuh "$here"/*.h "$root/include/pcoe.h" "$here/build.sh" 2>/dev/null | head -1)"
for src in "$here"/*.cu; do
  obj="$here/build/$(basename "${src%.cu}").o"
  objs+=("$obj")
  if [[ ! -f "$obj" || "$src" -nt "$obj" || "$newest_hdr" -nt "$obj" ]]; then
    "$NVCC" "${FLAGS[@]}" -c "$src" -o "$obj" &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [[ -n "$p" ]] && wait "$p"; done
"$NVCC" -shared -o "$out" "${objs[@]}"
echo "built $out"
