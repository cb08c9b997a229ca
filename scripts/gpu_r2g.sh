#!/bin/bash
mkdir -p gpurun_out
bash scripts/gpu_ab.sh "$@"
for lib in "$@"; do
TL_LIB_PATH=$PWD/dilabhelmholtzoct_b200/$lib timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'ph_small' -c 3 --csv --log-file gpurun_out/ab_$lib.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(l for l in open("gpurun_out/ab_$lib.csv") if l.startswith('"'))]
h=rows[0]; im,iv=h.index("Metric Name"),h.index("Metric Value")
print("$lib", [(r[im].split("__")[1][:12], r[iv]) for r in rows[1:7]])
PY
done
