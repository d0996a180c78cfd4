#!/bin/bash
# builds tuning variants of the library (same-spin kernel constants) under build/variants/, on the CPU
# usage: scripts/tune_ssg.sh NAME "-DF2_SSG_TW=512 -DF2_SSG_PFB=4"
set -e
NAME=$1; DEFS=$2
cd "$(dirname "$0")/../forte2_b200/csrc"
mkdir -p ../../tune_variants/$NAME
make -j4 OBJDIR=../../tune_variants/$NAME OUT=../../tune_variants/libf2_$NAME.so \
  NVFLAGS="-O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xptxas -v $DEFS" > ../../tune_variants/$NAME.log 2>&1
ls -la ../../tune_variants/libf2_$NAME.so
