#!/bin/bash
# dev: time the 128^3 / 160^3 / 256^2 apply on every package variant under variants/ (and the tree itself)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in base $(ls variants 2>/dev/null); do
  root=$PWD; [ $v != base ] && root=$PWD/variants/$v
  for prob in ${PROBS:-p3d128 cd3d160 p2d256}; do
    echo "== $v $prob"; ILUB_PKG_ROOT=$root timeout 300 python scripts/dev_one.py $prob 20 2>&1 | tail -2
  done
done 2>&1 | tee gpurun_out/sweep_variants.log
