"""Summarises a bench.py --layer-table CSV: per-layer and per-group kernel time per frame."""
import collections
import csv
import sys

path = sys.argv[1]
frames = float(sys.argv[2]) if len(sys.argv) > 2 else 1280.0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
rows = list(csv.reader(open(path)))
conv = [r for r in rows[1:] if r[0] == "conv"]
aux = [r for r in rows[1:] if r[0] == "aux"]
tot = sum(float(r[11]) for r in conv) + sum(float(r[11]) for r in aux)
print("total %.1f us/frame" % (1000 * tot / frames))
g = collections.defaultdict(float)
for r in conv:
    net = "bdcn" if (r[1].startswith("features") or r[1].startswith("msblock")) else "esf"
    g[(net, int(r[2]))] += float(r[11])
for k in sorted(g):
    print("  %-5s H=%-4d %7.1f us/frame" % (k[0], k[1], 1000 * g[k] / frames))
print("  conv total %.1f us/frame" % (1000 * sum(g.values()) / frames))
for r in aux:
    print("  %-18s %7.1f us/frame" % (r[1], 1000 * float(r[11]) / frames))
conv.sort(key=lambda r: -float(r[11]))
for r in conv[:top]:
    print("%-28s H%3s W%3s kpad%4s cout%4s taps%3s ntile%4s cfg%4s us/frame(all passes)%7.2f TF%7.1f" % (
        r[1], r[2], r[3], r[4], r[5], r[6], r[7], r[8], 1000 * float(r[11]) / frames, float(r[13])))
