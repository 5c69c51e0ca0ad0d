"""Turn an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table.
usage: python scripts/launch_table.py launches.csv [--per-launch] > profiles/rNN_<what>_launches.txt"""
import collections, csv, re, sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
ki, vi, gi, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Metric Name")
def short(n):
    if "at::" in n or "cub::" in n: return "torch:" + re.sub(r".*?(\w+_kernel|\w+Kernel).*", r"\1", n)[:40]
    n = re.sub(r"^void ", "", n); n = re.sub(r"\(.*", "", n); n = n.replace("einconv::", "")
    return n[:70]
agg = collections.OrderedDict(); total = 0.0
for r in rows:
    if r[mi] != "gpu__time_duration.sum": continue
    us = float(r[vi]) / 1e3; n = short(r[ki]); total += us
    if "--per-launch" in sys.argv: print(f"{us:10.1f} us  {r[gi]:>16s}  {n}")
    a = agg.setdefault(n, [0.0, 0]); a[0] += us; a[1] += 1
print(f"{'kernel':72s} {'launches':>8s} {'total us':>10s} {'share':>7s}")
for n, (us, k) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{n:72s} {k:8d} {us:10.1f} {us / total:7.1%}")
print(f"{'sum':72s} {'':8s} {total:10.1f}")
