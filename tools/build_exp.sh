#!/usr/bin/env bash
# Timing experiments (results are WRONG by construction): variants of libpcoe.so with parts of the v4 kernels
# compiled out, to see what bounds them.  usage: tools/build_exp.sh NAME -DPCOE_EXP_...   -> gpurun_exp/libpcoe_NAME.so
set -euo pipefail
root="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
src="$root/3d-pointcloud-orientation-estimation_b200/csrc"
name="$1"; shift
mkdir -p "$root/gpurun_exp"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Wno-deprecated-gpu-targets \
     "$@" -I"$root/include" -I"$src" -c "$src/sa.cu" -o "$root/gpurun_exp/sa_$name.o"
objs=()
for f in "$src"/build/*.o; do [[ "$(basename "$f")" == "sa.o" ]] || objs+=("$f"); done
nvcc -shared -o "$root/gpurun_exp/libpcoe_$name.so" "${objs[@]}" "$root/gpurun_exp/sa_$name.o"
rm -f "$root/gpurun_exp/sa_$name.o"
echo "built gpurun_exp/libpcoe_$name.so"
