"""Per-kernel DRAM traffic of one forward from an ncu CSV (dram__bytes_read.sum, dram__bytes_write.sum,
gpu__time_duration.sum; collected with --cache-control none so that L2 reuse BETWEEN kernels is kept):
    python tools/dram_traffic.py gpurun_out/dram_a.csv [skip_launches]"""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hi]; ix = {k: i for i, k in enumerate(h)}
per = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) < len(h): continue
    kid = int(r[ix["ID"]])
    if kid < skip: continue
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("void ", "").replace("ddb::", "")
    d = per.setdefault(kid, dict(name=name))
    v = float(r[ix["Metric Value"]]); u = r[ix["Metric Unit"]]
    m = r[ix["Metric Name"]]
    if m.startswith("dram__bytes"):
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    d[m] = v
agg = collections.OrderedDict()
for d in per.values():
    a = agg.setdefault(d["name"], [0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += d.get("dram__bytes_read.sum", 0); a[2] += d.get("dram__bytes_write.sum", 0); a[3] += d.get("gpu__time_duration.sum", 0)
tr = sum(a[1] for a in agg.values()); tw = sum(a[2] for a in agg.values())
print(f"# {sys.argv[1]}: {sum(a[0] for a in agg.values())} launches, DRAM read {tr/1e9:.2f} GB, write {tw/1e9:.2f} GB")
print(f"{'kernel':52s} {'n':>4s} {'read MB/launch':>15s} {'write MB/launch':>16s}")
for k, a in sorted(agg.items(), key=lambda kv: -(kv[1][1] + kv[1][2])):
    print(f"{k:52s} {a[0]:4d} {a[1]/a[0]/1e6:15.1f} {a[2]/a[0]/1e6:16.1f}")
