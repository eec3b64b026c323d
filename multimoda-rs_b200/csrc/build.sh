#!/usr/bin/env bash
# Builds libmmrs_b200.so (sm_100a only) in-tree. Usage: csrc/build.sh [extra nvcc flags]
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OUT=../libmmrs_b200.so
SRCS="mmrs_sweep.cu mmrs_host.cpp"
$NVCC -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
      -Xcompiler -fPIC,-O2,-Wall,-ffp-contract=off -Xptxas -v \
      --shared -x cu $SRCS -o $OUT "$@" 2> build.log || { cat build.log; exit 1; }
grep -E "error|warning" build.log | grep -v "ptxas info" | head -20 || true
echo "built $OUT"
