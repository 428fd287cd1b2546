#!/bin/bash
# Build a variant of libfhvae_b200.so with extra -D flags for A/B experiments on the GPU box:
#   tools/build_variant.sh NAME "-DWAVE_FOO=1 ..."   ->  build_variants/lib_NAME.so
# Objects of the unchanged translation units are cached in build_variants/obj (build_variants/ is git-ignored
# via *.so / *.o; it travels with gpurun snapshots).
set -e
cd "$(dirname "$0")/.."
NAME=$1; shift
DEFS="$*"
VFILES=${VFILES:-lstm_wave}     # translation units compiled with DEFS (space separated); the rest come from the cache
CS=pytorch_scalablefhvae_b200/csrc
OBJ=build_variants/obj
mkdir -p $OBJ
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
objs=""
for f in api gemm_simt gemm_tc gemm_wgrad gemm_proj lstm_simt lstm_cluster lstm_wave elbo heads disc table_adam_misc; do
  o=$OBJ/$f.o
  if [[ " $VFILES " == *" $f "* ]]; then
    o=$OBJ/${f}_$NAME.o
    nvcc $FLAGS $DEFS -c -o $o $CS/$f.cu &
    objs="$objs $o"
    continue
  fi
  if [ ! -f $o ] || [ $CS/$f.cu -nt $o ] || [ $CS/common.cuh -nt $o ] || [ $CS/tc_common.cuh -nt $o ] || [ include/fhvae_b200.h -nt $o ]; then
    nvcc $FLAGS -c -o $o $CS/$f.cu &
  fi
  objs="$objs $o"
done
wait
nvcc -shared -o build_variants/lib_$NAME.so $objs -lcudart
echo built build_variants/lib_$NAME.so "($DEFS)"
