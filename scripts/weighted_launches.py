"""Per-kernel GPU time of a ResNet-50 factor sweep from an ncu launch list of scripts/one_layer_factors.py
(one KFC + one KFAC-reduce call per distinct geometry), weighted by the layer counts of SURVEY.md App. C.
usage: python scripts/weighted_launches.py launches.csv"""
import collections, csv, re, sys
sys.path.insert(0, ".")
from bench import RESNET50_LAYERS
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
ki, vi, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
seq, cur = [], None
for r in rows:
    if r[mi] != "gpu__time_duration.sum": continue
    n = r[ki]
    if "distribution" in n:   # torch.rand of the next layer's input
        cur = []; seq.append(cur); continue
    if cur is None: continue
    n = re.sub(r"^void ", "", n); n = re.sub(r"\(.*", "", n).replace("einconv::", "").replace("tma::", "")
    cur.append((n[:44], float(r[vi]) / 1e3))
agg, tot = collections.defaultdict(float), 0.0
per_layer = []
for (c, h, k, s, p, count), ks in zip(RESNET50_LAYERS, seq):
    t = sum(us for _, us in ks)
    per_layer.append((c, h, k, s, count, t))
    for n, us in ks:
        agg[n] += us * count; tot += us * count
print(f"{'kernel':46s} {'us per sweep':>12s} {'share':>7s}")
for n, us in sorted(agg.items(), key=lambda kv: -kv[1]): print(f"{n:46s} {us:12.1f} {us / tot:7.1%}")
print(f"{'sum (GPU time of the kernels of one sweep)':46s} {tot:12.1f}")
print()
for c, h, k, s, count, t in per_layer: print(f"C{c:<5d} H{h:<4d} k{k} s{s} x{count}: {t:8.1f} us per layer (KFC + KFAC-reduce)")
