#!/bin/bash
# Build an A/B variant of the library: tools/build_variant.sh <name> <file.cu> [-DKNOB=value ...]
#   -> coldrec_b200/csrc/variants/lib_<name>.so (the named source recompiled with the flags, every other object as built by make)
# Run a probe against it with CR_LIB_PATH=coldrec_b200/csrc/variants/lib_<name>.so
set -e
cd "$(dirname "$0")/../coldrec_b200/csrc"
name=$1; src=$2; shift 2
mkdir -p variants
make -s -j8 >/dev/null
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-fvisibility=hidden -Xptxas -v --expt-relaxed-constexpr \
    "$@" -c "$src" -o "variants/${name}_${src%.cu}.o" 2> "variants/${name}.ptxas.log"
objs=""
for o in abi spmm score_simt score_tc score_api metrics towers tower_tc rowtopk train; do
    if [ "$o.cu" == "$src" ]; then objs="$objs variants/${name}_$o.o"; else objs="$objs $o.o"; fi
done
nvcc -shared -gencode arch=compute_100a,code=sm_100a -cudart static -o "variants/lib_${name}.so" $objs
grep -A1 "rows_grouped\|sweep_tc" "variants/${name}.ptxas.log" | grep "Used" | sort | uniq -c | head -4
