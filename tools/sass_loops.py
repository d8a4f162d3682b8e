"""Static issue-schedule view of a kernel's loops from its SASS control words (no GPU needed): for every backward branch,
the loop body's instruction count, the sum of the per-instruction stall counts (= minimum issue cycles of one warp for one
iteration, scoreboard waits excluded) and its instruction mix.
    python tools/sass_loops.py <kernel-name-substring> [--dump LO HI]   (addresses in hex: print that range)"""
import re
import subprocess
import sys
from collections import Counter
from pathlib import Path

so = Path(__file__).resolve().parents[1] / "duodiff_b200" / "libduodiff_b200.so"
pat = sys.argv[1]
out = subprocess.run(["cuobjdump", "-sass", str(so)], capture_output=True, text=True).stdout
blocks = out.split("Function : ")
for blk in blocks[1:]:
    name = blk.split("\n", 1)[0]
    dem = subprocess.run(["c++filt", name.strip()], capture_output=True, text=True).stdout.strip()
    if pat not in dem:
        continue
    lines = blk.splitlines()
    ins, i = [], 0
    while i < len(lines):
        m = re.match(r"\s*/\*([0-9a-f]{4,5})\*/\s+(.*?);\s*/\* (0x[0-9a-f]{16}) \*/", lines[i])
        if m and i + 1 < len(lines):
            m2 = re.match(r"\s*/\* (0x[0-9a-f]{16}) \*/", lines[i + 1])
            if m2:
                hi = int(m2.group(1), 16)
                ins.append((int(m.group(1), 16), m.group(2).strip(), (hi >> 41) & 0xF, (hi >> 52) & 0x3F))
                i += 2
                continue
        i += 1
    print(f"== {dem[:110]}  ({len(ins)} instructions)")
    if "--dump" in sys.argv:
        k = sys.argv.index("--dump")
        lo, hi = int(sys.argv[k + 1], 16), int(sys.argv[k + 2], 16)
        for x in ins:
            if lo <= x[0] <= hi:
                print(f"  {x[0]:5x} st={x[2]:2d} wm={x[3]:02x}  {x[1][:80]}")
        continue
    for x in ins:
        m = re.match(r"(@!?U?P\d+ )?BRA(\.U)? (0x[0-9a-f]+)", x[1])
        if not m:
            continue
        tgt = int(m.group(3), 16)
        if tgt >= x[0]:
            continue
        body = [y for y in ins if tgt <= y[0] <= x[0]]
        if len(body) < 24:
            continue
        ops = Counter()
        for y in body:
            w = y[1].split()
            op = (w[1] if w[0].startswith("@") and len(w) > 1 else w[0]).split(".")[0]
            ops[op] += 1
        print(f"  loop {tgt:5x}..{x[0]:5x}: {len(body):4d} instrs, stall sum {sum(y[2] for y in body):5d}, "
              f"branches {sum(1 for y in body if 'BRA' in y[1]):2d}  {dict(ops.most_common(9))}")
