#!/bin/bash
# Development aid: builds build/variants/lib_<name>.so = the current library with the evolution
# translation units of the listed chain lengths recompiled with extra -D flags (select the result
# with RC_LIB_PATH=...).
#   tools/build_variant.sh <name> <nspin[,nspin...]> [-DMACRO=value ...]
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
name=$1; ns=$2; shift 2
cd "$ROOT/code-robchar_b200/csrc"
odir="$ROOT/build/variants/obj_$name"
mkdir -p "$odir"
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --extended-lambda -Xcompiler -fPIC -diag-suppress 550,20091"
skip=""
for n in ${ns//,/ }; do
  ( nvcc $FLAGS -DRC_NSPIN=$n "$@" -Xptxas -v -c rc_fidelity_n.cu -o "$odir/rc_fidelity_$n.o" 2>&1 | grep -A2 "reg_kernelILi${n}ELi0ELb0" \
    | grep -o "Used [0-9]* registers\|[0-9]* bytes stack\|[0-9]* bytes spill stores" | tr '\n' ' '; echo " variant $name N=$n" ) &
  skip="$skip -e rc_fidelity_$n.o"
done
wait
nvcc -shared -o "$ROOT/build/variants/lib_$name.so" $(ls _obj/*.o | grep -v $skip) "$odir"/*.o -lcudart_static -lpthread -ldl -lrt 2>/dev/null
rm -rf "$odir"
