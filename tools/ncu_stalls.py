"""Top stalled SASS instructions of one kernel from `ncu -i rep --page source --csv` output."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
h = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[h]
ix = {k: i for i, k in enumerate(hdr)}
data = []
for r in rows[h + 1:]:
    if r and r[0] == "Kernel Name":
        break                      # next kernel of the report
    if len(r) == len(hdr) and r[0].startswith("0x"):
        data.append(r)
samp = "# Samples"
tot = sum(int(r[ix[samp]] or 0) for r in data)
stalls = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
agg = {k: sum(int(r[ix[k]] or 0) for r in data) for k in stalls}
print("total samples", tot)
print(", ".join("%s=%.1f%%" % (k[6:], 100.0 * v / max(tot, 1)) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
top = sorted(enumerate(data), key=lambda kv: -int(kv[1][ix[samp]] or 0))[:n]
for i, r in sorted(top):
    st = {k: int(r[ix[k]] or 0) for k in stalls}
    best = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print("%5d %6s %5.1f%% %-72s %s" % (i, r[ix[samp]], 100.0 * int(r[ix[samp]]) / max(tot, 1), r[ix["Source"]].strip()[:72],
                                       " ".join("%s:%d" % (k[6:], v) for k, v in best)))
