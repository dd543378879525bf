"""Per-source-line stall samples from `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv` (one launch)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr = None
data = []
for r in rows:
    if len(r) > 4 and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) == len(hdr) and r[0].isdigit():
        data.append(r)
iS = hdr.index("# Samples"); iE = hdr.index("Instructions Executed")
stall = {k: hdr.index(k) for k in ("stall_barrier", "stall_long_sb", "stall_short_sb", "stall_wait", "stall_mio",
                                   "stall_math", "stall_lg", "stall_not_selected", "stall_selected", "stall_branch_resolving")}
num = lambda s: int(s) if s.lstrip("-").isdigit() else 0
tot = sum(num(r[iS]) for r in data)
totE = sum(num(r[iE]) for r in data)
print(f"total samples {tot}, warp instructions {totE}")
print("line  samp%  inst%  | barrier long_sb short_sb wait mio math lg notsel selected branch | source")
for r in sorted(data, key=lambda r: -num(r[iS]))[:top]:
    s = num(r[iS])
    print(f"{r[0]:>4} {100*s/tot:6.1f} {100*num(r[iE])/totE:6.1f}  | " +
          " ".join(f"{100*num(r[i])/max(s,1):4.0f}" for i in stall.values()) + " | " + r[1].strip()[:90])
