#!/bin/bash
# runs tools/debug_one.py against differently compiled variants of the CUDA library
cp sahara_b200/libsahara_b200.so /tmp/orig.so
for v in build/var/*.so; do
  echo "== $v"
  cp $v sahara_b200/libsahara_b200.so
  python tools/debug_one.py 2>&1 | grep -E "DIFF|OK|mismatches" | sort | uniq -c
done
cp /tmp/orig.so sahara_b200/libsahara_b200.so
