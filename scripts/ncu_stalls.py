"""Summarise an `ncu --page source --csv` dump: total stall samples per reason and the top SASS lines."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_")]
tot = {h: 0 for h in stall_cols}
lines = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    try:
        samples = int(r[idx["# Samples"]])
    except ValueError:
        continue
    for h in stall_cols:
        try:
            tot[h] += int(r[idx[h]])
        except ValueError:
            pass
    lines.append((samples, r[idx["Source"]].strip(), {h: r[idx[h]] for h in stall_cols if r[idx[h]] not in ("0", "")}))
total = sum(s for s, _, _ in lines)
print("total samples", total)
for h, v in sorted(tot.items(), key=lambda kv: -kv[1])[:10]:
    print(f"  {h:28s} {v:8d}  {100.0 * v / max(total, 1):5.1f}%")
print("top lines:")
for s, src, st in sorted(lines, key=lambda t: -t[0])[: int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print(f"  {s:6d} {100.0 * s / max(total, 1):5.1f}%  {src[:70]:70s} {st}")
