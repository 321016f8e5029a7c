"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel
(UPD_* modes of k_update are told apart by the preceding kernel / stream order)."""
import csv, collections, re, sys

def main(path, title):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = [r for r in csv.DictReader(lines) if r.get("Metric Name") == "gpu__time_duration.sum"]
    agg = collections.OrderedDict()
    tot = 0.0
    for r in rows:
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("unnamed>::", "")
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        v = v / 1e6 if unit == "ns" else v / 1e3 if unit == "us" else v
        grid = int(r["Grid Size"].strip("()").split(",")[0])
        a = agg.setdefault(name, [0, 0.0, 0])
        a[0] += 1; a[1] += v; a[2] += grid
        tot += v
    print(f"# {title}\n")
    print(f"Launches: {len(rows)}; sum of kernel time {tot:.1f} ms (serialised, cold-cache: compare SHARES)\n")
    print("| kernel | launches | total ms | share | avg grid |\n|---|---:|---:|---:|---:|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {a[0]} | {a[1]:.2f} | {100*a[1]/tot:.1f}% | {a[2]//a[0]} |")

if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "ncu launch list")
