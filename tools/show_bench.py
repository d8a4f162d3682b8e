"""Compact view of a bench.py JSON line: python tools/show_bench.py gpurun_out/x.json"""
import json
import sys

for path in sys.argv[1:]:
    lines = [ln for ln in open(path).read().strip().splitlines() if ln.startswith("{")]
    if not lines:
        print(path, ": no JSON line")
        continue
    d = json.loads(lines[-1])
    print(f"== {path}")
    if "verify_shards" in d:
        print(json.dumps(d, indent=1))
        continue
    print(f"value {d.get('value')} {d.get('unit')}  e2e {d.get('e2e', {}).get('value')}  n_gpus {d.get('n_gpus')}  "
          f"ms/pass {d.get('ms_per_step')}  launches {d.get('gpu_launches')}  clocks {d.get('clocks')}")
    print("config:", d.get("config"))
    for key in ("step_ms_in_graph", "early_exit", "cpu_baseline"):
        if d.get(key):
            print(f"{key}: {d[key]}")
    r = d.get("roofline") or {}
    print("roofline:", {k: v for k, v in r.items() if k not in ("traffic_provenance", "peak_source")})
    for name, k in (d.get("kernels") or {}).items():
        print(f"  [{name}]")
        for c, v in k.items():
            print(f"     {c:14s} {v}")
