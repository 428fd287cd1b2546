#!/bin/bash
# Build a variant of libfhvae_b200.so with extra -D flags for A/B experiments on the GPU box:
#   tools/build_variant.sh NAME "-DWAVE_FOO=1 ..."   ->  build_variants/lib_NAME.so
# Objects of the unchanged translation units are cached in build_variants/obj (build_variants/ is git-ignored
# via *.so / *.o; it travels with gpurun snapshots).
set -e
cd "$(dirname "$0")/.."
NAME=$1; shift
DEFS="$*"
CS=pytorch_scalablefhvae_b200/csrc
OBJ=build_variants/obj
mkdir -p $OBJ
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
objs=""
for f in api gemm_simt gemm_tc gemm_wgrad gemm_proj lstm_simt lstm_cluster elbo heads disc table_adam_misc; do
  o=$OBJ/$f.o
  if [ ! -f $o ] || [ $CS/$f.cu -nt $o ] || [ $CS/common.cuh -nt $o ] || [ $CS/tc_common.cuh -nt $o ] || [ include/fhvae_b200.h -nt $o ]; then
    nvcc $FLAGS -c -o $o $CS/$f.cu &
  fi
  objs="$objs $o"
done
nvcc $FLAGS $DEFS -c -o $OBJ/lstm_wave_$NAME.o $CS/lstm_wave.cu &
wait
nvcc -shared -o build_variants/lib_$NAME.so $objs $OBJ/lstm_wave_$NAME.o -lcudart
echo built build_variants/lib_$NAME.so "($DEFS)"
