#!/bin/bash
# Build variants of the library that differ in the -D flags of one source file (kernel tuning experiments):
#   scripts/variant_libs.sh tech_stats.cu name1 "-DFOO=1" name2 "-DFOO=2" ...
# -> facet_b200/variants/lib_<name>.so ; select one with FACET_B200_LIB=<path>.
set -e
cd "$(dirname "$0")/.."
python -m facet_b200.build > /dev/null
src=$1; shift
mkdir -p facet_b200/variants
others=$(ls facet_b200/build/*.o | grep -v "/${src%.cu}.o")
while [ $# -gt 0 ]; do
  name=$1; flags=$2; shift 2
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr $flags \
       -c facet_b200/csrc/$src -o facet_b200/variants/${src%.cu}_$name.o
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -cudart static -o facet_b200/variants/lib_$name.so \
       facet_b200/variants/${src%.cu}_$name.o $others
  echo facet_b200/variants/lib_$name.so
done
