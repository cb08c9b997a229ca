"""Summarise an `ncu --metrics gpu__time_duration.sum,dram__bytes_* --csv` launch list per kernel."""
import collections, csv, sys
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ik, im, iv, iid = (hdr.index(x) for x in ("Kernel Name", "Metric Name", "Metric Value", "ID"))
per = collections.OrderedDict()
for r in rows[1:]:
    d = per.setdefault(r[iid], {"k": r[ik].split("(")[0][:60]})
    d[r[im]] = float(r[iv].replace(",", ""))
agg = collections.OrderedDict()
for d in per.values():
    a = agg.setdefault(d["k"], [0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += d.get("gpu__time_duration.sum", 0); a[2] += d.get("dram__bytes_read.sum", 0); a[3] += d.get("dram__bytes_write.sum", 0)
tot = sum(a[1] for a in agg.values())
unit = 1e6 if tot > 1e6 else 1e3  # ns or us
print("| kernel | launches | avg ms | share of listed time | DRAM read MB/launch | DRAM write MB/launch |")
print("|---|---|---|---|---|---|")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {a[0]} | {a[1]/a[0]/1e6:.3f} | {100*a[1]/tot:.1f}% | {a[2]/a[0]/1e6:.1f} | {a[3]/a[0]/1e6:.1f} |")
