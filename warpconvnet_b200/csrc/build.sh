#!/usr/bin/env bash
# Builds libwcn_b200.so in-tree for sm_100a. nvcc cross-compiles without a GPU.
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -diag-suppress 177"
# WCN_BRINGUP=1: the library honours the experiment switches of tools/exp_*.py (environment
# variables WCN_DEBUG / WCN_DEBUG_PTR / WCN_STAGES). The default build ignores them.
if [ "${WCN_BRINGUP:-0}" = "1" ]; then FLAGS="$FLAGS -DWCN_BRINGUP"; fi
# WCN_KERNEL_COUNTERS=1: compile the per-role cycle counters into the GEMM kernels (bring-up only)
if [ "${WCN_KERNEL_COUNTERS:-0}" = "1" ]; then FLAGS="$FLAGS -DWCN_KERNEL_COUNTERS -DWCN_BRINGUP"; fi
# WCN_ZERO_ROWS_SECOND_PASS=1: EXPERIMENT, off by default and not yet measured (profiles/r1i):
# missing neighbours of a gather stage are written with st.shared after the copies are issued
if [ "${WCN_ZERO_ROWS_SECOND_PASS:-0}" = "1" ]; then FLAGS="$FLAGS -DWCN_ZERO_ROWS_SECOND_PASS"; fi
# WCN_ENABLE_PDL=1: launch every kernel with the programmatic-stream-serialization attribute
# (off by default, see common.cuh)
if [ "${WCN_ENABLE_PDL:-0}" = "1" ]; then FLAGS="$FLAGS -DWCN_ENABLE_PDL"; fi
# WCN_EXTRA_FLAGS: extra nvcc flags (bring-up experiments, e.g. -DWCN_RN_UF=8)
FLAGS="$FLAGS ${WCN_EXTRA_FLAGS:-}"
mkdir -p build
pids=()
for f in cuhash coords conv_fwd conv_wgrad weight_prep knn rownorm conv_depthwise peer_allreduce capi; do
  $NVCC $FLAGS -c $f.cu -o build/$f.o &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o libwcn_b200.so build/cuhash.o build/coords.o build/conv_fwd.o build/conv_wgrad.o build/weight_prep.o build/knn.o build/rownorm.o build/conv_depthwise.o build/peer_allreduce.o build/capi.o
echo "built $(pwd)/libwcn_b200.so"
