#!/bin/bash
# Build a variant libudb with extra -D flags for same-box A/B runs:
#   tools/build_variant.sh /root/repo/variants/libudb_trace.so -DUDB_ATTN_TRACE
# then run with UDB_LIB=<that path>.
set -e
out=$1; shift
cd "$(dirname "$0")/../unidepth_b200/csrc"
mkdir -p "$(dirname "$out")" /tmp/udbvar
for f in common gemm conv_halo attention elementwise v1_kernels engine engine_v1 p2p; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC \
     --expt-relaxed-constexpr "$@" -c $f.cu -o /tmp/udbvar/$f.o &
done
wait
/usr/local/cuda/bin/nvcc -shared -o "$out" /tmp/udbvar/*.o -lcudart
echo "$out"
