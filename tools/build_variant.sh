#!/bin/bash
# build_variant.sh NAME [GIT_REV] [nvcc -D flags...] -> build/variants/NAME/{libsahara_b200.so,libsahara_host.so}
# A/B builds of the CUDA library for tools/ab_probe.py (SB200_LIB_DIR selects the directory at run time).
# GIT_REV "-" = the working tree; otherwise the csrc/ and include/ of that revision.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
NAME=$1; REV=${2:--}; shift; shift || true
OUT=$ROOT/build/variants/$NAME
mkdir -p "$OUT"
SRC=$ROOT
if [ "$REV" != "-" ]; then
  SRC=$ROOT/build/scratch/src_$NAME
  rm -rf "$SRC"; mkdir -p "$SRC"
  git -C "$ROOT" archive "$REV" sahara_b200/csrc include | tar -x -C "$SRC"
fi
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC \
  -Xcompiler -Wno-deprecated-declarations -ccbin /usr/bin/g++ "$@" -shared "$SRC/sahara_b200/csrc/capi.cu" -o "$OUT/libsahara_b200.so"
cp "$ROOT/sahara_b200/libsahara_host.so" "$OUT/"
echo "built $OUT"
