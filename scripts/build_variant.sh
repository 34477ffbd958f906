#!/bin/bash
# build a tuning variant of libspgpu.so: scripts/build_variant.sh NAME -DSPG_RB=128 -DSPG_MINB=4 ...
set -e
name=$1; shift
mkdir -p build/var/$name
for f in spartan_parallel_b200/csrc/*.cu; do
  b=$(basename $f .cu)
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -c $f -o build/var/$name/$b.o &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/var/libspgpu_$name.so build/var/$name/*.o -lcudart
rm -rf build/var/$name
echo built build/var/libspgpu_$name.so
