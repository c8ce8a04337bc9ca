"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: share / total / count / average per kernel."""
import csv, sys, re, collections
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
agg = collections.OrderedDict()
for r in rows[1:]:
    if r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = r[ix["Kernel Name"]]
    name = re.sub(r"\(.*", "", name)
    v = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    v = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)   # -> us
    a = agg.setdefault(name, [0.0, 0])
    a[0] += v; a[1] += 1
tot = sum(a[0] for a in agg.values())
print(f"sum of kernel durations: {tot/1e3:.1f} ms over {sum(a[1] for a in agg.values())} launches")
print("| share | total ms | launches | avg us | kernel |\n|---|---|---|---|---|")
for name, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[: int(sys.argv[2]) if len(sys.argv) > 2 else 20]:
    print(f"| {100*a[0]/tot:.1f} % | {a[0]/1e3:.2f} | {a[1]} | {a[0]/a[1]:.1f} | `{name[:90]}` |")
