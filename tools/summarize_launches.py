"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections, csv, re, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
tot = collections.defaultdict(float); cnt = collections.Counter()
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    v = float(row["Metric Value"].replace(",", "")) * {"ns": 1, "us": 1e3, "ms": 1e6}.get(row["Metric Unit"], 1)
    tot[name] += v; cnt[name] += 1
T = sum(tot.values())
print(f"total kernel time {T/1e6:.3f} ms over {sum(cnt.values())} launches")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{v/1e6:9.3f} ms {100*v/T:5.1f}% n={cnt[k]:5d} avg {v/cnt[k]/1e3:8.1f} us  {k[:90]}")
