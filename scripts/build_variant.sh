#!/bin/bash
# Build a variant of libbtrc_b200.so with extra nvcc flags for device/kernels.cu into btrc_b200/variants/libbtrc_<name>.so
# (A/B runs select it with BTRC_B200_LIB).  usage: scripts/build_variant.sh <name> "<nvcc flags>"
set -e
NAME=$1; EXTRA=$2
cd "$(dirname "$0")/.."
python -m btrc_b200.build > /dev/null
mkdir -p btrc_b200/variants btrc_b200/build/variants
O=btrc_b200/build/variants/kernels_$NAME.o
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -ccbin /usr/bin/g++ -Xcompiler -fPIC --expt-relaxed-constexpr \
  -ftz=true -prec-div=false -prec-sqrt=false $EXTRA -c btrc_b200/csrc/device/kernels.cu -o $O
OBJS=$(ls btrc_b200/build/*.o | grep -v device_kernels.cu.o)
/usr/local/cuda/bin/nvcc -shared -cudart static -ccbin /usr/bin/g++ -o btrc_b200/variants/libbtrc_$NAME.so $OBJS $O -Xcompiler -fopenmp -lz -lgomp -ldl
/usr/local/cuda/bin/cuobjdump -res-usage $O | grep -A1 "shade_kernelILi2ELb0" | grep -o "REG:[0-9]* STACK:[0-9]*"
