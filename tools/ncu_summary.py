"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time share per kernel (and grid)."""
import collections, csv, re, sys

def main(path, top=40):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)
        name = re.sub(r"\(.*", "", row["Kernel Name"])[:64]
        key = f"{name} grid={row['Grid Size']}"
        agg[key][0] += 1
        agg[key][1] += v
    tot = sum(v[1] for v in agg.values())
    n = sum(v[0] for v in agg.values())
    print(f"# {path}: {n} launches, {tot:.1f} us total kernel time (serialised, cold cache: compare SHARES)")
    print("# share   total_us   launches   avg_us   kernel")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{t / tot * 100:6.2f}% {t:10.1f} {c:6d} {t / c:9.1f}   {k}")
    byk = collections.defaultdict(float)
    for k, (c, t) in agg.items():
        byk[k.split(" grid=")[0]] += t
    print("# by kernel (all grids)")
    for k, t in sorted(byk.items(), key=lambda kv: -kv[1])[:20]:
        print(f"{t / tot * 100:6.2f}% {t:10.1f}   {k}")

if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
