"""Development tool: per-instruction execution counts / stall samples from `ncu --page source --csv`.
usage: dev_ncu_src.py file.csv [min_exec]  -> prints instructions executed more than min_exec times,
grouped in address order, with stall samples."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
base = int(data[0][ix["Address"]], 16)
min_exec = int(sys.argv[2]) if len(sys.argv) > 2 else 0
tot_samples = sum(int(r[ix["# Samples"]] or 0) for r in data)
tot_exec = sum(int(r[ix["Instructions Executed"]] or 0) for r in data)
print(f"total samples {tot_samples}, total warp-instructions executed {tot_exec}")
for r in data:
    ex = int(r[ix["Instructions Executed"]] or 0)
    if ex >= min_exec:
        print(f"{int(r[ix['Address']], 16) - base:#7x} {ex:9d} {int(r[ix['# Samples']] or 0):6d}  {r[ix['Source']].strip()[:90]}")
