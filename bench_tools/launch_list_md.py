"""Markdown summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list:  python bench_tools/launch_list_md.py in.csv "title" > out.md"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hi]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1e-3)
    name = re.sub(r"\(.*", "", r[ki]).strip()
    tot[name] += v
    cnt[name] += 1
total = sum(tot.values())
ours = {k: v for k, v in tot.items() if "naqs" in k or "stats_" in k}
print(f"# ncu launch list — {sys.argv[2]}\n")
print("`ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 200` — cold-cache, serialised launches: compare SHARES, not absolutes.\n")
print("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|")
for k, v in tot.most_common():
    print(f"| `{k}` | {cnt[k]} | {v:.1f} | {v / cnt[k]:.1f} | {100 * v / total:.1f}% |")
hot = max(ours, key=ours.get)
print(f"\nShare of the hot kernel (`{hot}`) among this library's kernels: {100 * ours[hot] / sum(ours.values()):.1f}% "
      f"(the torch fill kernel is the L2 flush between steps).")
