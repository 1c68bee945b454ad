"""Summarise an ncu source-page CSV (ncu -i X.ncu-rep --page source --csv): top SASS instructions by stall
samples, instruction mix by opcode, and stall-reason totals."""
import collections
import csv
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[2:]:           # first profiled launch only (the file repeats the table per launch)
    if r and r[0] == "Kernel Name":
        break
    if len(r) == len(hdr):
        data.append(r)
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
reasons = collections.Counter()
opmix = collections.Counter()
opsamples = collections.Counter()
for r in data:
    ins = r[ix["Source"]].strip()
    op = ins.split()[0] if ins else "?"
    if op.startswith("@"):
        op = ins.split()[1]
    op = op.split(".")[0]
    opmix[op] += int(r[ix["Instructions Executed"]] or 0)
    opsamples[op] += int(r[ix["# Samples"]] or 0)
    for c in stall_cols:
        reasons[c] += int(r[ix[c]] or 0)
print(f"total samples {tot}")
print("stall reasons:", ", ".join(f"{k[6:]} {100*v/max(1,sum(reasons.values())):.1f}%" for k, v in reasons.most_common(8)))
tinst = sum(opmix.values())
print("instruction mix (warp-level executed):", ", ".join(f"{k} {100*v/tinst:.1f}%" for k, v in opmix.most_common(14)))
print("samples by opcode:", ", ".join(f"{k} {100*v/max(1,tot):.1f}%" for k, v in opsamples.most_common(12)))
print("top instructions by samples:")
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:top]:
    s = int(r[ix["# Samples"]] or 0)
    why = max(stall_cols, key=lambda c: int(r[ix[c]] or 0))
    print(f"  {100*s/max(1,tot):5.1f}%  {r[ix['Source']].strip()[:70]:70s}  {why[6:]}")
