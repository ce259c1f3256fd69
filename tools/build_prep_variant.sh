#!/bin/bash
# usage: tools/build_prep_variant.sh <name> [-DFLAG ...]  ->  quantizedmha_b200/lib/variants/libqmha_<name>.so
# Variant of the library with extra preprocessor flags on prepare.cu (quantiser A/B runs: QMHA_LIB selects the library).
set -e
name=$1; shift
cd "$(dirname "$0")/.."
mkdir -p build/var quantizedmha_b200/lib/variants
GEN="-gencode arch=compute_100a,code=sm_100a"
nvcc -O3 -std=c++17 -lineinfo $GEN -Xcompiler -fPIC "$@" -c quantizedmha_b200/csrc/prepare.cu -o build/var/prepare_$name.o
nvcc -shared $GEN -o quantizedmha_b200/lib/variants/libqmha_$name.so build/obj/attn_fwd.o build/var/prepare_$name.o build/obj/api_fa_tc_int8_b.o
echo "built quantizedmha_b200/lib/variants/libqmha_$name.so"
