"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel."""
import csv, collections, re, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
per = collections.defaultdict(list)
for row in csv.DictReader(lines):
    v = float(row["Metric Value"]); u = row["Metric Unit"]
    v = v / 1000 if u == "ns" else v * 1000 if u == "ms" else v
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("vlg::", "")
    per[name + " grid" + row.get("Grid Size", "")].append(v)
tot = sum(map(sum, per.values()))
print(f"total {tot:.1f} us over {sum(map(len, per.values()))} launches")
for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k[:78]:78s} n={len(v):4d} sum={sum(v):8.1f}us avg={sum(v)/len(v):7.2f} share={100*sum(v)/tot:5.1f}%")
