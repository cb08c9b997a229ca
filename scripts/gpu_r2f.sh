#!/bin/bash
# A/B of gradient-job variants (libtopoloss_<v>.so): tail timeline of the fused launch
for v in "$@"; do
  echo "== $v"
  TL_LIB_PATH=$PWD/dilabhelmholtzoct_b200/libtopoloss_$v.so timeout 120 python scripts/tail_probe.py 2>&1 | head -6
done
