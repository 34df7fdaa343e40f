"""Development tool: per-kernel hot instructions of an `ncu --page source --csv` export (SASS view).
usage: dev_ncu_top.py file.csv kernel_substring [top_n]"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
pat = sys.argv[2]
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
# split into kernels
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
for b in blocks:
    if pat not in b["name"]:
        continue
    hdr = b["rows"][0]
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in b["rows"][1:] if len(r) == len(hdr)]
    base = int(data[0][ix["Address"]], 16)
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = Counter()
    for r in data:
        for h in stall_cols:
            tot[h] += int(r[ix[h]] or 0)
    n_s = sum(int(r[ix["# Samples"]] or 0) for r in data)
    n_e = sum(int(r[ix["Instructions Executed"]] or 0) for r in data)
    print("=====", b["name"][:90])
    print(f"samples {n_s}  warp-instructions {n_e}")
    print("  " + "  ".join(f"{h[6:]}={c}" for h, c in tot.most_common(10)))
    # opcode histogram weighted by samples
    ops = Counter(); opsx = Counter()
    for r in data:
        t = r[ix["Source"]].strip()
        parts = t.split()
        op = parts[1] if parts and parts[0].startswith("@") else (parts[0] if parts else "?")
        op = op.split(".")[0]
        ops[op] += int(r[ix["# Samples"]] or 0)
        opsx[op] += int(r[ix["Instructions Executed"]] or 0)
    print("  by opcode (samples / executed):")
    for op, c in ops.most_common(18):
        print(f"    {op:10s} {c:7d} {opsx[op]:10d}")
    order = sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:top_n]
    for r in sorted(order, key=lambda r: int(r[ix["Address"]], 16)):
        st = {h[6:]: int(r[ix[h]] or 0) for h in stall_cols}
        top = ", ".join(f"{k}={v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3] if v)
        print(f"{int(r[ix['Address']], 16) - base:#7x} ex={int(r[ix['Instructions Executed']] or 0):8d} s={int(r[ix['# Samples']] or 0):5d}  {r[ix['Source']].strip()[:70]:70s} {top}")
