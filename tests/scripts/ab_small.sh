#!/bin/bash
cd "$(dirname "$0")/../.."
for so in ptz-calib_b200/csrc/ab/*.so; do
  cp "$so" ptz-calib_b200/csrc/libptzcalib_b200.so
  echo "== $so"
  python tests/scripts/small_breakdown.py 2>&1 | grep -E "cfg|pcg"
done
