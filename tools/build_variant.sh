#!/bin/bash
# build_variant.sh NAME "EXTRA nvcc flags": a developer build of libdilqr (DILQR_FAST_BUILD:
# the bench / smoke shapes only) into differentiable-ilqr_b200/variants/libdilqr_NAME.so,
# for A/B measurements:  DILQR_LIB=.../libdilqr_NAME.so python bench.py ...
set -e
NAME=$1; EXTRA=$2
ROOT=$(cd "$(dirname "$0")/.." && pwd)
SRC=$ROOT/differentiable-ilqr_b200/csrc
OBJ=/tmp/dilqr_var_$NAME; mkdir -p $OBJ $ROOT/differentiable-ilqr_b200/variants
NVCC=/usr/local/cuda/bin/nvcc
FLAGS="-O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC -gencode arch=compute_100a,code=sm_100a -DDILQR_FAST_BUILD $EXTRA"
cd $SRC
for t in 0 1; do for g in 0 1 2 3; do
  $NVCC $FLAGS -DDILQR_SCALAR_F64=$t -DDILQR_GROUP=$g -c api.cu -o $OBJ/api_${t}_${g}.o &
done; done
$NVCC $FLAGS -c dispatch.cu -o $OBJ/dispatch.o &
$NVCC $FLAGS -c cost_glue.cu -o $OBJ/cost_glue.o &
wait
$NVCC -shared -gencode arch=compute_100a,code=sm_100a -o $ROOT/differentiable-ilqr_b200/variants/libdilqr_$NAME.so $OBJ/*.o
echo built $NAME
