"""Per-kernel totals of an ncu launch list (--metrics gpu__time_duration.sum --csv).  usage: launch_summary.py file.csv[.gz] [steps]"""
import collections, csv, gzip, io, re, sys

path = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 0
f = io.TextIOWrapper(gzip.open(path)) if path.endswith(".gz") else open(path)
rows = list(csv.reader(f))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]
kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
d = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    name = r[kn]
    m = re.search(r"vrr::(?:\(anonymous namespace\)::|<unnamed>::)?(\w+)(<[^(]{0,70})?", name)
    key = (m.group(1) + (m.group(2) or "")) if m else name[:70]
    t = float(r[mv].replace(",", ""))
    t = t / 1e3 if r[mu] == "ns" else t * 1e3 if r[mu] == "ms" else t
    d[key][0] += 1
    d[key][1] += t
tot = sum(v[1] for v in d.values())
for k, v in sorted(d.items(), key=lambda kv: -kv[1][1])[:32]:
    per = f" per-step {v[1] / steps / 1e3:6.2f} ms" if steps else ""
    print(f"{k[:88]:88s} n={v[0]:5d} tot={v[1] / 1e3:7.2f} ms avg={v[1] / v[0]:7.1f} us {100 * v[1] / tot:5.1f}%{per}")
print(f"total {tot / 1e3:.2f} ms over {sum(v[0] for v in d.values())} launches")
