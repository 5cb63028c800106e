"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel name the launch count, total and
average duration and the share of the listed time.  Usage: python tools/launch_list_summary.py in.csv out.txt [header]"""
import csv
import re
import sys

src, dst = sys.argv[1], sys.argv[2]
header = sys.argv[3] if len(sys.argv) > 3 else ""
rows = [r for r in csv.reader(l for l in open(src, errors="replace") if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
ni, mi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}       # -> us
agg = {}
for r in rows:
    name = re.sub(r"\(.*", "", r[ni]).replace("void ", "").replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    t = float(r[mi].replace(",", "")) * scale.get(r[ui], 1.0)
    a = agg.setdefault(name[:90], [0, 0.0])
    a[0] += 1; a[1] += t
tot = sum(a[1] for a in agg.values())
with open(dst, "w") as f:
    if header:
        f.write("# " + header + "\n")
    f.write(f"# {len(rows)} launches, {tot / 1e3:.2f} ms listed; gpu__time_duration.sum, --clock-control none; per-launch times are cold-cache and\n"
            "# serialised: compare SHARES, not absolutes\n")
    for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        f.write(f"{t / 1e3:9.2f} ms {c:6d}x  avg {t / c:8.1f} us  share {100 * t / tot:5.1f}%  {n}\n")
print(open(dst).read()[:3000])
