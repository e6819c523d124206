"""Summarise an `ncu --page source --csv` export of a warp-specialised kernel: per role region (split at USETMAXREG)
samples, instructions executed, top stall reasons, and the hottest instructions.  usage: ncu_src_summary.py src.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 12
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
regions, cur = [], []
for r in data:
    cur.append(r)
    if "USETMAXREG" in r[ix["Source"]]:
        regions.append(cur); cur = []
regions.append(cur)
# region 0 = prologue + producer preamble up to first SETMAXREG; the SETMAXREG ends the *previous* role's preamble
tot = sum(int(r[ix["# Samples"]]) for r in data)
print("total samples", tot, "instructions executed", sum(int(r[ix["Instructions Executed"]]) for r in data))
for n, reg in enumerate(regions):
    s = sum(int(r[ix["# Samples"]]) for r in reg); ie = sum(int(r[ix["Instructions Executed"]]) for r in reg)
    st = sorted(((sum(int(r[ix[k]]) for r in reg), k) for k in stalls), reverse=True)[:6]
    print(f"--- region {n}: {len(reg)} instrs, samples {s} ({100*s/max(tot,1):.1f}%), executed {ie}")
    print("    stalls:", ", ".join(f"{k[6:]} {v}" for v, k in st if v))
    hot = sorted(reg, key=lambda r: -int(r[ix["# Samples"]]))[:top]
    for r in hot:
        rs = sorted(((int(r[ix[k]]), k[6:]) for k in stalls), reverse=True)[:2]
        print(f"    {int(r[ix['# Samples']]):6d} {int(r[ix['Instructions Executed']]):9d}  {r[ix['Source']].strip()[:60]:60s} {rs}")
