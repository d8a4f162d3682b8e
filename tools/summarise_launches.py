"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
    python tools/summarise_launches.py gpurun_out/launches_r01.csv > profiles/r01_launches_summary.txt"""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hi]; ix = {k: i for i, k in enumerate(h)}
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) < len(h) or r[ix["Metric Name"]] != "gpu__time_duration.sum": continue
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("void ", "").replace("ddb::", "")
    v = float(r[ix["Metric Value"]]); u = r[ix["Metric Unit"]]
    v = v / 1e3 if u in ("nsecond", "ns") else v  # -> us
    agg.setdefault(name, []).append(v)
tot = sum(sum(v) for v in agg.values())
print(f"# {sys.argv[1]}: {sum(len(v) for v in agg.values())} launches, {tot:.0f} us total (cold-cache, serialised: compare shares)")
print(f"{'kernel':60s} {'n':>5s} {'mean us':>9s} {'share':>7s}")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:60s} {len(v):5d} {sum(v)/len(v):9.1f} {100*sum(v)/tot:6.1f}%")
