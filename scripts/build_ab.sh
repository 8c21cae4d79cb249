#!/bin/bash
# A/B builds of libsdr_b200.so with extra -D flags: bash scripts/build_ab.sh <name> "<flags>"  ->  rtl-sdr-rs_b200/lib/ab/libsdr_<name>.so
# (select at run time with SDR_B200_LIB=<path>; see rtl-sdr-rs_b200/_ffi.py)
set -e
NAME=$1; FLAGS=$2
cd "$(dirname "$0")/../rtl-sdr-rs_b200"
mkdir -p build/ab_$NAME lib/ab
ARCH="-gencode arch=compute_100a,code=sm_100a"
for f in csrc/*.cu; do
  nvcc $ARCH -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr $FLAGS -c $f -o build/ab_$NAME/$(basename $f .cu).o &
done
for f in csrc/*.cpp; do
  nvcc -O2 -std=c++17 -Ibuild -Xcompiler -fPIC $FLAGS -c $f -o build/ab_$NAME/$(basename $f .cpp).o &
done
wait
nvcc $ARCH -shared -o lib/ab/libsdr_$NAME.so build/ab_$NAME/*.o -lpthread -ldl
echo built lib/ab/libsdr_$NAME.so
