#!/bin/bash
# tools/ab/build_variant.sh <name> [extra nvcc flags]: another build of the current sources as tools/ab/librtz_<name>.so (A/B on one GPU box via RTZ_LIB)
set -e
cd "$(dirname "$0")/../.."
name=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC -shared "$@" -o tools/ab/librtz_$name.so raytracing-with-zig_b200/csrc/rtz_api.cu
