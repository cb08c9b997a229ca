"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export per source line.
usage: python scripts/ncu_lines.py report.ncu-rep [top_n] [lo hi]   (lo/hi: line range filter of ph_small.cuh)"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
lo = int(sys.argv[3]) if len(sys.argv) > 3 else 0; hi = int(sys.argv[4]) if len(sys.argv) > 4 else 10**9
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
per = []; cur_file = None; hdr = None
tot_i = tot_s = 0
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] in ("Function Name",): continue
    if r[0] == "Line No": hdr = r; continue
    if r[0] != "": per.append({"f": cur_file, "l": int(r[0]), "src": r[1], "i": 0, "s": 0, "t": 0, "st": {}}); continue
    if len(r) < 9 or not per: continue
    try: samp = int(r[6]); inst = int(r[7]); tinst = int(r[8])
    except ValueError: continue
    p = per[-1]; p["i"] += inst; p["s"] += samp; p["t"] += tinst; tot_i += inst; tot_s += samp
    for name, v in zip(hdr[32:49], r[32:49]):
        try: v = int(v)
        except ValueError: continue
        if v: p["st"][name] = p["st"].get(name, 0) + v
print("total warp-instructions", tot_i, "samples", tot_s)
sel = [p for p in per if not (p["f"] == "ph_small.cuh" and not (lo <= p["l"] <= hi))]
if lo: sel = [p for p in per if p["f"] == "ph_small.cuh" and lo <= p["l"] <= hi]
print("selected: inst %.1f%% samples %.1f%%" % (100 * sum(p["i"] for p in sel) / tot_i, 100 * sum(p["s"] for p in sel) / tot_s))
sel.sort(key=lambda p: -p["s"])
for p in sel[:top]:
    st = sorted(p["st"].items(), key=lambda kv: -kv[1])[:3]
    print(f"{p['f']}:{p['l']:4d} inst={100*p['i']/tot_i:4.1f}% samp={100*p['s']/tot_s:4.1f}% thr={p['t']/max(1,p['i']):4.1f} {' '.join(k.replace('stall_','')+':'+str(v) for k,v in st):40s} | {p['src'].strip()[:80]}")
