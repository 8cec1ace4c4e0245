#!/bin/bash
# Builds an A/B variant of the C-ABI library with extra -D flags: tools/build_variant.sh <suffix> <nvcc flags...>
# -> tfg-pathtracer_b200/csrc/libeleven_b200_<suffix>.so (git-ignored; select it with ELEVEN_LIB=<path>).
set -e
cd "$(dirname "$0")/../tfg-pathtracer_b200/csrc"
sfx=$1; shift
nvcc -O3 -std=c++17 -lineinfo --fmad=false -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-O3 "$@" -c eleven_api.cu -o /tmp/eleven_api_$sfx.o
g++ -O3 -std=c++17 -fPIC -pthread "$@" -c bvh8_build.cpp -o /tmp/bvh8_build_$sfx.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o libeleven_b200_$sfx.so /tmp/eleven_api_$sfx.o /tmp/bvh8_build_$sfx.o -lpthread
echo built libeleven_b200_$sfx.so
