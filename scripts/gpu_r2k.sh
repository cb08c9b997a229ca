#!/bin/bash
# unroll factors of the fast front end's trip loops (variant libraries) against the new default (TL_HOPS=4)
mkdir -p gpurun_out
echo "== default"; AB_MODES=0 timeout 100 python scripts/ab_list_mode.py 2>&1 | grep "^mode\|fingerprint" | cut -c1-100
for v in "$@"; do
  echo "== $v"; TL_LIB_PATH=$PWD/dilabhelmholtzoct_b200/libtopoloss_$v.so AB_MODES=0 timeout 100 python scripts/ab_list_mode.py 2>&1 | grep "^mode\|fingerprint" | cut -c1-100
done
