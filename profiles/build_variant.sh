#!/bin/bash
# build the current csrc/ into profiles/_variants/libmacm_<name>.so (A/B measurements through MACM_LIB)
set -e
cd "$(dirname "$0")/.."
SRC=${2:-gym-macm_b200/csrc}
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -prec-div=true -prec-sqrt=true -ftz=false \
  $EXTRA -Xcompiler -fPIC -shared -I include -I $SRC -o profiles/_variants/libmacm_$1.so $SRC/macm_kernels.cu $SRC/macm_kernels_huge.cu $SRC/macm_aux.cu $SRC/macm_api.cu
