#!/bin/bash
# measurement build of the library: bash tools/build_variant.sh <name> "<-D...>"  ->  bokego_b200/libbokego_b200_<name>.so
# (only csrc/bk_forward.cu -- or the source named by a third argument, e.g. bk_train_tc -- is recompiled; the other objects are
# those of the last regular build)
set -e
cd "$(dirname "$0")/.."
n=$1; d=$2; f=${3:-bk_forward}
o=/tmp/${f}_$n.o
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -diag-suppress 1886 $d -c bokego_b200/csrc/$f.cu -o $o
objs=$(ls bokego_b200/csrc/*.o | grep -v $f.o)
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o bokego_b200/libbokego_b200_$n.so $o $objs -lcudart
echo built $n
