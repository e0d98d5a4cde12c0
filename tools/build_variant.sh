#!/bin/bash
# Build a tuning variant of the kernel library: tools/build_variant.sh NAME -DFLAG=... -> pegainfer_b200/variants/libNAME.so
# (git-ignored; select it with ModelRuntimeConfig(kernel_lib=...) or tools/quick_decode.py --lib).
set -e
name=$1; shift
cd "$(dirname "$0")/.."
out=pegainfer_b200/variants; obj=pegainfer_b200/build/var_$name
mkdir -p $out $obj
for f in pegainfer_b200/csrc/*.cu; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 --std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr \
    -ccbin /usr/bin/g++ "$@" -c $f -o $obj/$(basename ${f%.cu}).o &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/lib$name.so $obj/*.o -lcudart -ccbin /usr/bin/g++
echo built $out/lib$name.so
