#!/bin/bash
# crossing-edge list in L2: timing + bit-equality per mode, then DRAM bytes per mode (2 launches each)
mkdir -p gpurun_out
timeout 200 python scripts/ab_list_mode.py > gpurun_out/ab_list_mode.log 2>&1; echo "timing rc=$?"; cat gpurun_out/ab_list_mode.log | grep -v Warn
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:ph_small -c 10 --csv --log-file gpurun_out/ab_list_mode_ncu.csv python scripts/ab_list_mode.py ncu > gpurun_out/ab_list_mode_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'P'
import csv
rows = list(csv.reader(l for l in open("gpurun_out/ab_list_mode_ncu.csv") if l.startswith('"')))
h = rows[0]; iid, im, iv = h.index("ID"), h.index("Metric Name"), h.index("Metric Value")
per = {}
for r in rows[1:]:
    per.setdefault(int(r[iid]), {})[r[im]] = float(r[iv].replace(",", ""))
for k in sorted(per):
    d = per[k]
    print(k, {a.split("__")[1][:14]: round(v / 1e6, 2) for a, v in d.items()})
P
# compile-time merge knobs (variant libraries), mode 0 only
for v in hops2 hops4 refill4 refill16; do
  echo "== $v"; TL_LIB_PATH=$PWD/dilabhelmholtzoct_b200/libtopoloss_$v.so AB_MODES=0 timeout 100 python scripts/ab_list_mode.py 2>&1 | grep "^mode"
done
