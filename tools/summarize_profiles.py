"""Turns the raw ncu CSVs of tools/profile_round.sh (gpurun_out/<round>_*.csv) into the tracked
summaries under profiles/: a launch-share table, per-layer DRAM traffic of the convolution kernel
(+ the JSON bench.py reads for roofline.traffic) and key metrics of the full captures.

    python tools/summarize_profiles.py r01
"""
import csv
import json
import os
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = sys.argv[1] if len(sys.argv) > 1 else "r01"
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")

BDCN_ORDER = []
_v = ["conv1_1", "conv1_2", "conv2_1", "conv2_2", "conv3_1", "conv3_2", "conv3_3", "conv4_1", "conv4_2", "conv4_3", "conv5_1", "conv5_2", "conv5_3"]
_m = ["1_1", "1_2", "2_1", "2_2", "3_1", "3_2", "3_3", "4_1", "4_2", "4_3", "5_1", "5_2", "5_3"]
_merged = {0, 2, 4, 5} if R >= "r02" else set()      # round 2: msblock i's conv rides with features.conv(i+1) (engine.cuh build_bdcn)
for i in range(13):
    if i > 0 and (i - 1) not in _merged:
        BDCN_ORDER.append("features." + _v[i])
    if i in _merged:
        BDCN_ORDER.append("features.%s+msblock%s.conv" % (_v[i + 1], _m[i]))
    else:
        BDCN_ORDER.append("msblock%s.conv" % _m[i])
    BDCN_ORDER.append("msblock%s.tail" % _m[i])
FRAMES = int(os.environ.get("PROFILE_FRAMES", "16"))
ESF_ORDER = ["enc.head.conv2"]
for b in ["down_block1", "down_block2", "down_block3", "down_block4", "bottleneck"]:
    ESF_ORDER += ["enc.%s.%s" % (b, c) for c in ["conv1", "conv21", "conv22", "conv31", "conv32", "TD.conv"]]
for b in ["up_block4", "up_block3", "up_block2", "up_block1"]:
    ESF_ORDER += ["dec.%s.%s" % (b, c) for c in ["pre", "conv11", "conv12", "conv21", "conv22"]]
ESF_ORDER += ["dec.final.conv1", "dec.final.conv2", "elReg.c1"]


def read_ncu_csv(path):
    rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
    hdr = rows[0]
    return hdr, rows[1:]


def launches():
    hdr, rows = read_ncu_csv(os.path.join(G, R + "_launches.csv"))
    iname, ival = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows:
        try:
            v = float(r[ival].replace(",", ""))
        except Exception:
            continue
        n = r[iname].split("(")[0]
        if n.startswith("void "):
            n = n[5:]
        if n.startswith("conv_tc_kernel"):          # <false> plain / <true> upsample-add epilogue instantiations
            n = "conv_tc_kernel"
        agg[n][0] += 1
        agg[n][1] += v
    tot = sum(v[1] for v in agg.values())
    out = ["# %s: the first 960 kernel launches (warm-up step + both timed steps) of `bench.py --steps 2 --warmup 1` under" % R,
           "# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: shares, not absolutes)",
           "kernel,launches,total_us,share"]
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("%s,%d,%.1f,%.4f" % (n, c, t / 1000.0, t / tot))
    open(os.path.join(P, R + "_launch_shares.csv"), "w").write("\n".join(out) + "\n")
    return agg, tot


def conv_dram():
    res = []
    for net, order in (("bdcn", BDCN_ORDER), ("esf", ESF_ORDER)):
        hdr, rows = read_ncu_csv(os.path.join(G, "%s_conv_dram_%s.csv" % (R, net)))
        iid, imn, ival, iun = hdr.index("ID"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
        per = defaultdict(dict)
        for r in rows:
            try:
                v = float(r[ival].replace(",", ""))
            except Exception:
                continue
            u = r[iun]
            if u == "Mbyte": v *= 1e6
            elif u == "Kbyte": v *= 1e3
            elif u == "Gbyte": v *= 1e9
            elif u == "us": v *= 1e3
            elif u == "ms": v *= 1e6
            per[int(r[iid])][r[imn]] = v
        for k, (lid, m) in enumerate(sorted(per.items())):
            name = order[k] if k < len(order) else "?"
            res.append((net, name, m.get("dram__bytes_read.sum", 0), m.get("dram__bytes_write.sum", 0), m.get("gpu__time_duration.sum", 0)))
    out = ["# %s: DRAM traffic and duration of every conv_tc_kernel launch of one warm %d-frame pass (ESF encoder: %d frames)" % (R, FRAMES, 2 * FRAMES),
           "net,layer,dram_read_MB,dram_write_MB,duration_us,dram_GBps"]
    tb = tt = 0
    for net, name, rd, wr, ns in res:
        out.append("%s,%s,%.2f,%.2f,%.1f,%.0f" % (net, name, rd / 1e6, wr / 1e6, ns / 1e3, (rd + wr) / max(ns, 1)))
        tb += rd + wr
        tt += ns
    sfx = "" if FRAMES == 16 else "_b%d" % FRAMES
    open(os.path.join(P, R + "_conv_dram%s.csv" % sfx), "w").write("\n".join(out) + "\n")
    js = {"dram_bytes_per_launch": tb / len(res), "launches": len(res), "frames": FRAMES, "config": "baseline_edge",
          "note": "mean dram__bytes_read.sum + dram__bytes_write.sum over the %d conv_tc_kernel launches of one warm %d-frame "
                  "baseline_edge pass (ncu, profiles/%s_conv_dram%s.csv)" % (len(res), FRAMES, R, sfx),
          "total_dram_bytes_per_frame": tb / FRAMES, "total_conv_us_per_frame_under_ncu": tt / 1e3 / FRAMES}
    json.dump(js, open(os.path.join(P, R + "_conv_traffic%s.json" % sfx), "w"), indent=1)
    return js


WANT = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.avg"]


def full(name, labels):
    path = os.path.join(G, "%s_%s_raw.csv" % (R, name))
    if not os.path.isfile(path):
        return []
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    out = []
    for k, r in enumerate(rows[2:]):
        d = {"capture": name, "launch": labels[k] if k < len(labels) else str(k), "kernel": r[ix["Kernel Name"]].split("(")[0]}
        for w in WANT:
            if w in ix:
                d[w] = r[ix[w]] + " " + units[ix[w]]
        out.append(d)
    return out


def main():
    os.makedirs(P, exist_ok=True)
    agg, tot = launches()
    js = conv_dram()
    caps = []
    if R >= "r02":
        caps += full("conv_bdcn", ["features.conv1_2+msblock1_1.conv", "msblock1_1.tail (phase lattice)", "msblock1_2.conv", "msblock1_2.tail (phase lattice)",
                                   "features.conv2_1", "features.conv2_2+msblock2_1.conv"])
        caps += full("conv_vgg3_2", ["features.conv3_2+msblock3_1.conv"])
        caps += full("conv_dec", ["dec.up_block1.pre", "dec.up_block1.conv11 (upsample-add)", "dec.up_block1.conv12", "dec.up_block1.conv21 (upsample-add)"])
    else:
        caps += full("conv_bdcn", ["msblock1_1.conv", "msblock1_1.tail", "features.conv1_2", "msblock1_2.conv", "msblock1_2.tail", "features.conv2_1"])
        caps += full("conv_vgg3_2", ["features.conv3_2"])
    caps += full("conv_esf", ["enc.head.conv2", "enc.down_block1.conv1", "enc.down_block1.conv21", "enc.down_block1.conv22",
                              "enc.down_block1.conv31", "enc.down_block1.conv32", "enc.down_block1.TD.conv"])
    caps += full("aux_esf", [])
    caps += full("aux_bdcn", [])
    lines = ["# %s: key metrics of the `ncu --set full --clock-control none` captures (16-frame micro-batch; ESF encoder kernels see 32 frames)" % R, ""]
    for d in caps:
        lines.append("## %s / %s (%s)" % (d["capture"], d["launch"], d["kernel"]))
        for w in WANT:
            if w in d:
                lines.append("- %s: %s" % (w, d[w]))
        lines.append("")
    open(os.path.join(P, R + "_ncu_full_summary.md"), "w").write("\n".join(lines))
    print("conv share of launches: %.3f" % (agg["conv_tc_kernel"][1] / tot))
    print(json.dumps(js))


main()
