#!/bin/bash
# (The variant macros of round 1 — CMOE_LINE_BATCH, CMOE_SPLIT_CHAINS, CMOE_EXP_TABLE, CMOE_EXP_GUARD, CMOE_PMAX_FP,
# CMOE_COV_ROWS/UNROLL/MINB — were resolved to the shipped configuration after the experiments; the script stays as the
# harness for -D experiments: add an #ifndef/#define default in the source, then build one library per flag set.)
# Builds experiment variants of the fused q-KG kernel (only kg_mc_inst_8.cu is recompiled) into variants/libvar_<name>.so;
# bench.py picks one up through CMOE_B200_LIB.  Usage: [FILE=cov] profiles/build_variants.sh name "-DFLAG=1 ..." [name flags]...
# FILE = the translation unit to recompile (default kg_mc_inst_8).
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
P=$ROOT/cornell-moe_b200
mkdir -p $ROOT/variants
make -C $P -j8 lib >/dev/null
FILE=${FILE:-kg_mc_inst_8}
OTHERS=$(ls $P/build/*.o | grep -v /$FILE.o)
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -std=c++17 -O3 -lineinfo \
    -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr $flags -c $P/csrc/$FILE.cu -o $ROOT/variants/var_$name.o \
    2> $ROOT/variants/var_$name.ptxas.log
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -shared -o $ROOT/variants/libvar_$name.so \
    $OTHERS $ROOT/variants/var_$name.o -cudart static
  grep -A2 "${ENTRY:-kg_mc_kernelILi0ELi8ELi8}" $ROOT/variants/var_$name.ptxas.log | grep -o "Used [0-9]* registers\|[0-9]* bytes spill stores" | tr '\n' ' '
  echo " <- $name ($flags)"
done
