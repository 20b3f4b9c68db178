#!/bin/bash
# Builds librobir_b200.so in-tree (sm_100a only).  Called by __graft_entry__.build().
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
ARCH="-gencode arch=compute_100a,code=sm_100a"
COMMON="-O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=default --expt-relaxed-constexpr"
mkdir -p build
obj() { # src out extra
  if [ ! -f "$2" ] || [ "$1" -nt "$2" ] || [ -n "$(find . -maxdepth 1 \( -name '*.h' -o -name '*.cuh' \) -newer "$2")" ]; then
    echo "nvcc $1"
    rm -f "$2"                      # a failed compile must not leave a stale object for the link step
    $NVCC $ARCH $COMMON $3 -c "$1" -o "$2"
  fi
}
pids=()
obj capi.cu build/capi.o "" & pids+=($!)
obj vis.cu build/vis.o "-Xptxas -v" & pids+=($!)
obj sg.cu build/sg.o "" & pids+=($!)
obj sdf.cu build/sdf.o "" & pids+=($!)
obj trace.cu build/trace.o "-fmad=false" & pids+=($!)
obj vis_tc.cu build/vis_tc.o "-Xptxas -v" & pids+=($!)
obj mlp.cu build/mlp.o "-Xptxas -v" & pids+=($!)
obj sphere_trace.cu build/sphere_trace.o "-Xptxas -v" & pids+=($!)
obj loss.cu build/loss.o "" & pids+=($!)
obj tc_mlp.cu build/tc_mlp.o "-Xptxas -v" & pids+=($!)
obj neus.cu build/neus.o "" & pids+=($!)
obj sdf_tc.cu build/sdf_tc.o "-Xptxas -v" & pids+=($!)
for pid in "${pids[@]}"; do wait "$pid" || { echo "build.sh: a compile job failed" >&2; exit 1; }; done
$NVCC $ARCH -shared -o ../librobir_b200.so build/capi.o build/vis.o build/sg.o build/sdf.o build/trace.o build/vis_tc.o build/mlp.o build/sphere_trace.o build/loss.o build/tc_mlp.o build/neus.o build/sdf_tc.o -lcudart_static -lpthread -ldl -lrt
echo "built $(cd .. && pwd)/librobir_b200.so"
