"""SASS evidence: per kernel, counts of the memory / atomic / vote instructions that carry the design.
usage: python scripts/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections, os, re, subprocess
so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dilabhelmholtzoct_b200", "libtopoloss.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
pat = re.compile(r"\b(LDG\.[A-Z0-9.]+|STG\.[A-Z0-9.]+|LDS(?:\.[A-Z0-9.]+)?|STS(?:\.[A-Z0-9.]+)?|ATOMS\.[A-Z0-9.]+|ATOMG\.[A-Z0-9.]+|ATOM\.[A-Z0-9.]+|RED\.[A-Z0-9.]+|LDGSTS\.[A-Z0-9.]+|UBLKCP\.[A-Z0-9.]+|SYNCS\.[A-Z0-9.]+|UTMA[A-Z0-9.]*|VOTE\.[A-Z]+|REDUX\.[A-Z]+|SHFL\.[A-Z]+|BAR\.SYNC[A-Z_.]*|MEMBAR\.[A-Z.]+)\b")
per = collections.OrderedDict(); cur = None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = per.setdefault(m.group(1), collections.Counter()); continue
    if cur is None: continue
    m = pat.search(line)
    if m: cur[m.group(1)] += 1
print("# SASS evidence (cuobjdump -sass dilabhelmholtzoct_b200/libtopoloss.so), instruction counts per kernel")
print("# ATOMS.CAS.64 / .128 : packed / wide triplet-table entries in shared memory;  ATOMG.E.CAS.* : global spill / global kernel")
print("# LDGSTS.E.BYPASS.128 : cp.async staging of the crossing-edge list;  LDS.64 : packed table hops and 4-entry union-find reads;")
print("# UBLKCP.S.G / UBLKCP.G.S + SYNCS.* : bulk asynchronous copies (cp.async.bulk) with mbarrier completion -- pair records in, gradient tile out (tail's gradient jobs)")
print("# LDG.E.128.CONSTANT : 128-bit map loads (4 pixels per lane);  STG.E.128 : 128-bit zero fill / records;  no HMMA / UTCMMA anywhere")
for k, c in per.items():
    print("==", k)
    for name, n in sorted(c.items(), key=lambda kv: -kv[1]):
        print(f"{n:7d} {name}")
tc = len(re.findall(r"\b(HMMA|UTCHMMA|UTCMMA|IMMA|QMMA)", txt))
print("== tensor-core instructions in the whole library:", tc)
