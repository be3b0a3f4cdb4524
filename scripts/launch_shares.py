"""Summarise an `ncu --csv --metrics gpu__time_duration.sum` log: share and launch count per kernel name."""
import collections
import csv
import sys

src, dst = sys.argv[1], sys.argv[2]
lines = [l for l in open(src) if l.startswith('"')]
rows = list(csv.reader(lines))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[1:]:
    if len(r) < len(hdr):
        continue
    try:
        v = float(r[ix["Metric Value"]].replace(",", ""))
    except ValueError:
        continue
    unit = r[ix["Metric Unit"]]
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
    name = r[ix["Kernel Name"]]
    name = name.split("(")[0] if not name.startswith("void") else name[5:].split("(")[0]
    tot[name[-90:]] += v
    cnt[name[-90:]] += 1
s = sum(tot.values())
with open(dst, "w") as f:
    f.write(f"total {s:.1f} us over {sum(cnt.values())} launches\n")
    for k, v in tot.most_common(40):
        f.write(f"{100 * v / s:5.1f}%  {v:9.1f} us  x{cnt[k]:4d}  {k}\n")
print(open(dst).read())
