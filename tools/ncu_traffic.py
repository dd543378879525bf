"""profiles/traffic.json from `ncu -i X.ncu-rep --page raw --csv`: per-launch dram__bytes_read.sum + dram__bytes_write.sum
of a kernel at a known shape (bench.py reports it as roofline.traffic for exactly that shape, never scaled).
    python tools/ncu_traffic.py raw.csv KERNEL_SUBSTR M K1p SRC_NOTE"""
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw, kern, M, K1p, src = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
rows = list(csv.reader(l for l in open(raw) if not l.startswith("==")))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


vals = []
for r in rows[2:]:
    if kern in r[ix["Kernel Name"]]:
        vals.append(sum(to_bytes(r[ix[m]], units[ix[m]]) for m in ("dram__bytes_read.sum", "dram__bytes_write.sum")))
assert vals, "kernel not found in the report"
path = os.path.join(ROOT, "profiles", "traffic.json")
db = json.load(open(path)) if os.path.exists(path) else {}
ent = {"M": M, "K1p": K1p, "dram_bytes": sum(vals) / len(vals), "launches": len(vals), "src": src}
db[kern] = [e for e in db.get(kern, []) if not (e["M"] == M and e["K1p"] == K1p)] + [ent]
json.dump(db, open(path, "w"), indent=1)
print(ent)
