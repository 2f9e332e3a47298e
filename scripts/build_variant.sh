#!/usr/bin/env bash
# build_variant.sh <name> <file.cu> "<extra nvcc flags>": lib/libgaitb200_<name>.so with ONE source recompiled with extra flags
set -eu
NAME=$1; SRC=$2; EXTRA=$3
PKG=video-based-gait-analysis-for-dementia_b200
mkdir -p build/var_$NAME
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Iinclude -I$PKG/csrc --cudart static $EXTRA -c $PKG/csrc/$SRC.cu -o build/var_$NAME/$SRC.o
OBJS=$(ls build/*.o | grep -v "/$SRC.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared --cudart static -o $PKG/lib/libgaitb200_$NAME.so $OBJS build/var_$NAME/$SRC.o
echo built $PKG/lib/libgaitb200_$NAME.so
