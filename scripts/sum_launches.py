"""Summarise an ncu --csv launch list (gpu__time_duration.sum) by kernel name."""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]; kn = h.index("Kernel Name"); mv = h.index("Metric Value"); mu = h.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    try: v = float(r[mv].replace(",", ""))
    except ValueError: continue
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[mu], 1e-3)
    a = agg.setdefault(r[kn][:90], [0, 0.0]); a[0] += 1; a[1] += v * scale
tot = sum(a[1] for a in agg.values())
print(f"total {tot/1e3:.3f} ms over {sum(a[0] for a in agg.values())} launches")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print(f"{t/1e3:9.3f} ms {100*t/tot:5.1f}%  n={n:4d} avg={t/n:8.1f} us  {k}")
