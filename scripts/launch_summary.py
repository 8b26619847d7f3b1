"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections, csv, re, sys
path, title = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else ""
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rows = list(csv.DictReader(lines))
tot, cnt = collections.defaultdict(float), collections.Counter()
for row in rows:
    name = re.sub(r"\(.*", "", row["Kernel Name"]); name = re.sub(r"^void tmesh::", "", name)
    v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
    v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
    tot[name] += v; cnt[name] += 1
T = sum(tot.values())
print(title)
print("launch-serialised, cold-cache times (ncu): compare shares, not absolutes")
print(f"{len(rows)} launches, {T:.0f} us of kernel time\n")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{v:10.1f} us {100*v/T:5.1f}% n={cnt[k]:5d} avg={v/cnt[k]:8.2f} us  {k}")
