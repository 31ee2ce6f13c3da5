#!/bin/bash
# Builds a tuning variant of libm3dgpu.so: scripts/build_variant.sh NAME "-DFOO=1 ..." ["FILE.cu ..."]
# (FILE.cu ...: the translation units the flags apply to, default trace_kernels.cu)
# -> variants/libm3dgpu_NAME.so (select it with M3D_LIB=variants/libm3dgpu_NAME.so).
set -e
cd "$(dirname "$0")/../model3d_b200/csrc"
NAME=$1; FLAGS_EXTRA=$2; UNITS=${3:-trace_kernels.cu}
mkdir -p ../../variants build_$NAME
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-DM3D_HAVE_SCENE -DM3D_HAVE_RAYCAST -DM3D_HAVE_PATH -DM3D_HAVE_BIDIR -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-Wall,-Wno-unused-function,-pthread --expt-relaxed-constexpr $FLAGS_EXTRA"
# only the named translation units differ; reuse the other objects of the main build
OBJS=$(ls build/*.o)
for UNIT in $UNITS; do
  OBJ=${UNIT%.cu}.o
  $NVCC $FLAGS -c $UNIT -o build_$NAME/$OBJ &
  OBJS=$(echo "$OBJS" | grep -v "/$OBJ")
done
wait
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o ../../variants/libm3dgpu_$NAME.so $OBJS build_$NAME/*.o -lcudart
rm -rf build_$NAME
echo built variants/libm3dgpu_$NAME.so
