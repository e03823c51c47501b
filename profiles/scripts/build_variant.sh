#!/bin/bash
# build_variant.sh NAME "-DFLAG=.. -DFLAG2=.." ["flags for b2j_entropy_fused.cu only" ["flags for b2j_fused.cu only" ["flags for b2j_entropy.cu only"]]]  ->  jpeg_encoder_gsoc2017_ishan-bansal_b200/variants/libb2j_NAME.so
# (a copy of the library with other compile-time knobs, for A/B timing on one GPU box; git-ignored)
set -e
NAME=$1; FLAGS=$2
# (the per-file ptxas levels of csrc/Makefile unless given, "" = ptxas' default)
EFLAGS=${3--Xptxas -regUsageLevel=8}; KFLAGS=${4--Xptxas -regUsageLevel=3}; SFLAGS=$5
ROOT=$(cd "$(dirname "$0")/../.." && pwd)
SRC=$ROOT/jpeg_encoder_gsoc2017_ishan-bansal_b200/csrc
OUT=$ROOT/jpeg_encoder_gsoc2017_ishan-bansal_b200/variants
mkdir -p $OUT
TMP=$(mktemp -d)
ARCH="-gencode arch=compute_100a,code=sm_100a"
for f in b2j_common b2j_stages b2j_fused b2j_entropy b2j_entropy_fused b2j_encode b2j_png; do
  X=""; [ $f = b2j_entropy_fused ] && X="$EFLAGS"; [ $f = b2j_fused ] && X="$KFLAGS"; [ $f = b2j_entropy ] && X="$SFLAGS"
  nvcc $ARCH -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC -Xptxas -v $FLAGS $X -c -o $TMP/$f.o $SRC/$f.cu 2> $TMP/$f.log &
done
wait
nvcc $ARCH -shared -o $OUT/libb2j_$NAME.so $TMP/*.o -lcudart
grep -A2 "k_fused_pairILi2ELi2E\|k_entropy_fusedILb1ELb0" $TMP/b2j_fused.log $TMP/b2j_entropy_fused.log | grep -E "registers|spill" | tr '\n' ' '; echo " -> libb2j_$NAME.so"
rm -rf $TMP
