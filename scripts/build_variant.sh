#!/bin/bash
# scripts/build_variant.sh NAME SOURCE.cu "-DFOO=1 -DBAR=2": a second libedgegs.so in which ONE source file is compiled with
# extra nvcc flags -> edgegaussians_b200/_C/libedgegs_NAME.so (select it at run time with EG_LIB=<that path>).
set -e
NAME=$1; SRC=$2; shift; shift
C=edgegaussians_b200/_C
python -m edgegaussians_b200.build > /dev/null
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr $* \
     -c edgegaussians_b200/csrc/$SRC -o $C/variant_$NAME.o
OBJS=$(ls $C/eg_*.o | grep -v "/${SRC%.cu}.o")
nvcc -shared -o $C/libedgegs_$NAME.so $OBJS $C/variant_$NAME.o -gencode arch=compute_100a,code=sm_100a -lcudart -ldl
rm -f $C/variant_$NAME.o
echo built $C/libedgegs_$NAME.so
