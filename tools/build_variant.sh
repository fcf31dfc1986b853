#!/bin/bash
# Builds an experimental variant of the library: tools/build_variant.sh NAME "EXTRA NVCC FLAGS" file1.cu [file2.cu ...]
# The named sources are recompiled with the extra flags, everything else is taken from csrc/build/; the result is
# pldepth_b200/variants/NAME.so (selected at run time through PLDEPTH_B200_LIB; git-ignored, travels with gpurun).
set -e
NAME=$1; FLAGS=$2; shift 2
cd "$(dirname "$0")/../pldepth_b200/csrc"
mkdir -p build/var_$NAME ../variants
OBJS=""
for f in pld_runtime pld_api pld_step pld_eval pld_sort pld_pilot pld_lists_small pld_lists_large pld_lists_tab pld_score_reg; do
  [ -f $f.cu ] || continue
  if [[ " $* " == *" $f.cu "* ]]; then
    /usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-Wall -Xptxas -v \
      --expt-relaxed-constexpr $FLAGS -c $f.cu -o build/var_$NAME/$f.o 2> build/var_$NAME/$f.ptxas.log || (cat build/var_$NAME/$f.ptxas.log; exit 1)
    OBJS="$OBJS build/var_$NAME/$f.o"
  else
    OBJS="$OBJS build/$f.o"
  fi
done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/$NAME.so $OBJS -lcudart
echo built pldepth_b200/variants/$NAME.so
