#!/bin/bash
# usage: tools/build_variant.sh <name> [-DFLAG ...]   ->  quantizedmha_b200/lib/variants/libqmha_<name>.so
# Builds a variant of the library with extra preprocessor flags on attn_fwd.cu (same-box A/B with tools/ab_libs.py).
set -e
name=$1; shift
cd "$(dirname "$0")/.."
make -s lib >/dev/null
mkdir -p build/var quantizedmha_b200/lib/variants
GEN="-gencode arch=compute_100a,code=sm_100a"
nvcc -O3 -std=c++17 -lineinfo $GEN -Xcompiler -fPIC --ptxas-options=-v "$@" -c quantizedmha_b200/csrc/attn_fwd.cu -o build/var/attn_fwd_$name.o 2> build/var/attn_fwd_$name.ptxas.log || { tail -20 build/var/attn_fwd_$name.ptxas.log; exit 1; }
nvcc -shared $GEN -o quantizedmha_b200/lib/variants/libqmha_$name.so build/var/attn_fwd_$name.o build/obj/prepare.o build/obj/api_fa_tc_int8_b.o
echo "built quantizedmha_b200/lib/variants/libqmha_$name.so"
