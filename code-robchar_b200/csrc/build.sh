#!/bin/bash
# Builds librobchar_b200.so for sm_100a (in-tree).  Usage: build.sh [jobs]
set -e
cd "$(dirname "$0")"
JOBS=${1:-8}
OUT=${RC_OUT:-../librobchar_b200.so}   # RC_OUT / RC_OBJ / RC_EXTRA_FLAGS: tuning builds next to the shipped one
OBJ=${RC_OBJ:-_obj}
mkdir -p $OBJ
NVCC=${NVCC:-nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --extended-lambda -Xcompiler -fPIC -diag-suppress 550,20091 ${RC_EXTRA_FLAGS:-}"
cmds=()
for n in $(seq 2 16); do
  cmds+=("$NVCC $FLAGS -DRC_NSPIN=$n -c rc_fidelity_n.cu -o $OBJ/rc_fidelity_$n.o")
done
for f in rc_fidelity rc_stats rc_rank rc_api rc_expm rc_peer rc_objective rc_grad rc_kendall_large rc_dense_mc; do
  cmds+=("$NVCC $FLAGS -c $f.cu -o $OBJ/$f.o")
done
printf '%s\n' "${cmds[@]}" | xargs -P "$JOBS" -I{} bash -c "{}"
$NVCC -shared -o $OUT $OBJ/*.o -lcudart_static -lpthread -ldl -lrt
echo "built $OUT"
