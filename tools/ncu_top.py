"""Summarise an .ncu-rep: key metrics, stall-reason totals and the top stalled instructions.
    python tools/ncu_top.py gpurun_out/prof.ncu-rep [N]"""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = {h: (v, u) for h, u, v in zip(hdr, units, r)}
    print("==", d["Kernel Name"][0][:90], d["Grid Size"][0], d["Block Size"][0])
    for k in ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
              "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
              "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
              "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
              "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
              "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread"]:
        if k in d: print(f"  {k:78s} {d[k][0]:>16s} {d[k][1]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if "Address" in r and "Source" in r)
hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}; data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
def num(x):
    try: return float(x)
    except ValueError: return 0.0
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(num(r[ix["# Samples"]]) for r in data)
agg = {s[6:]: int(sum(num(r[ix[s]]) for r in data)) for s in stalls}
print("total samples", int(tot), {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
ops = collections.Counter()
for r in data:
    w = r[ix["Source"]].split()
    if not w: continue
    op = (w[1] if w[0].startswith("@") and len(w) > 1 else w[0]).split(".")[0]
    ops[op] += num(r[ix["Instructions Executed"]])
print("instr mix (M warp-instr):", {k: round(v / 1e6, 2) for k, v in ops.most_common(14)})
for r in sorted(data, key=lambda r: -num(r[ix["# Samples"]]))[:N]:
    st = {s[6:]: int(num(r[ix[s]])) for s in stalls if num(r[ix[s]]) > 0}
    print(r[ix["Address"]][-5:], f"{int(num(r[ix['# Samples']])):6d} {int(num(r[ix['Instructions Executed']])):9d}", r[ix["Source"]][:70], st)
