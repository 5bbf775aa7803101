"""Groups a bench.py --layer-table CSV into the buckets VERDICT.md uses (us per real frame; the shared
encoder runs on 2 frames per real frame when add_edge, which `frames` in the table already counts)."""
import collections
import csv
import sys

path = sys.argv[1]
real_frames = float(sys.argv[2])          # steps * batch of the timed region
rows = list(csv.reader(open(path)))
g = collections.OrderedDict()


def add(k, ms):
    g[k] = g.get(k, 0.0) + ms


for r in rows:
    if r[0] not in ("conv", "aux"):
        continue
    name, ms = r[1], float(r[11])
    if r[0] == "aux":
        add("aux:" + name, ms)
        continue
    taps = int(r[6])
    if name.startswith("features"):
        add("VGG (+merged msblock.conv)", ms)
    elif name.endswith(".tail"):
        add("MSBlock tails", ms)
    elif name.startswith("msblock"):
        add("MSBlock .conv (standalone)", ms)
    elif name.startswith("dec.") and taps == 1 and not name.endswith(".pre"):
        add("ESF decoder 1x1 (+upsample-add)", ms)
    elif taps == 1:
        add("ESF 1x1 (encoder, pre)", ms)
    else:
        add("ESF 3x3 (+ head c1, logits)", ms)
tot = sum(g.values())
for k, v in g.items():
    print("%-36s %7.1f us/frame  %5.1f %%" % (k, 1000 * v / real_frames, 100 * v / tot))
print("%-36s %7.1f us/frame" % ("total (kernel time)", 1000 * tot / real_frames))
