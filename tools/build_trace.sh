#!/usr/bin/env bash
# Debug build with the in-kernel clock trace of the v4 kernels (-DPCOE_TC4_TRACE) -> libpcoe_trace.so
# Use:  PCOE_LIB=$PWD/3d-pointcloud-orientation-estimation_b200/libpcoe_trace.so python tools/trace_sa.py sa1 bwd
set -euo pipefail
root="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
src="$root/3d-pointcloud-orientation-estimation_b200/csrc"
mkdir -p "$src/build_trace"
objs=()
for f in "$src"/*.cu; do
  o="$src/build_trace/$(basename "${f%.cu}").o"; objs+=("$o")
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Wno-deprecated-gpu-targets \
       -DPCOE_TC4_TRACE -I"$root/include" -I"$src" -c "$f" -o "$o" &
done
wait
nvcc -shared -o "$root/3d-pointcloud-orientation-estimation_b200/libpcoe_trace.so" "${objs[@]}"
echo built libpcoe_trace.so
