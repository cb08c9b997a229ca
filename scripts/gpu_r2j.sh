#!/bin/bash
# merge-loop unroll (TL_HOPS) / refill threshold variants, and run-to-run identity of loss and gradient
mkdir -p gpurun_out
AB_MODES=0,0,1 timeout 100 python scripts/ab_list_mode.py 2>&1 | grep "^mode"
for v in hops4 hops5 hops6 hops8 hops4r4 hops4r12; do
  echo "== $v"; TL_LIB_PATH=$PWD/dilabhelmholtzoct_b200/libtopoloss_$v.so AB_MODES=0 timeout 100 python scripts/ab_list_mode.py 2>&1 | grep "^mode" | cut -c1-60
done
