#!/usr/bin/env bash
# Builds an experimental variant: build_variant.sh <name> <extra nvcc flags...>  -> ../variants/libmmrs_<name>.so
set -euo pipefail
cd "$(dirname "$0")"
name=$1; shift
mkdir -p ../variants
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
   -Xcompiler -fPIC,-O2,-ffp-contract=off --shared -x cu mmrs_sweep.cu mmrs_host.cpp -o ../variants/libmmrs_$name.so "$@" 2>/dev/null
echo built $name
