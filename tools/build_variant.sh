#!/bin/bash
# usage: tools/build_variant.sh NAME "-DFLAG=1 ..."  ->  paddlemix_b200/csrc/build/variants/libb200mix_NAME.so
# (A/B kernel experiments: run with B200MIX_LIB=<that path>)
set -e
cd "$(dirname "$0")/../paddlemix_b200/csrc"
name=$1; shift
out=build/variants; mkdir -p $out/$name
for f in common gemm attention norm elementwise collective; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC \
    --expt-relaxed-constexpr $@ -c $f.cu -o $out/$name/$f.o &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libb200mix_$name.so $out/$name/*.o -cudart static -ldl
echo "$(pwd)/$out/libb200mix_$name.so"
