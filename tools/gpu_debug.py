"""Layer-by-layer comparison of the engine against the CPU oracle on the golden B=2 input.

    python tools/gpu_debug.py simt|tc [config]

Prints one line per tensor (max abs error, reference magnitude) and never stops at the first
mismatch, so a single GPU run localises a bug.  Test infrastructure (imports oracle/)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
mode = sys.argv[1] if len(sys.argv) > 1 else "tc"
cfg = sys.argv[2] if len(sys.argv) > 2 else "baseline_edge"
os.environ["EGN_CONV"] = mode
os.environ["EGN_NO_ARENA"] = "1"     # intermediate maps are read back after the forward: no buffer sharing

import numpy as np
import torch
import torch.nn.functional as F

import egn_b200
from oracle import graph, synth

torch.set_num_threads(os.cpu_count() or 8)
dev = torch.device("cuda:0")
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
img = torch.from_numpy(np.load(os.path.join(root, "tests/golden/fwd_input.npz"))["img"])


def report(name, got, ref):
    got = np.asarray(got, np.float64); ref = np.asarray(ref, np.float64)
    if got.shape != ref.shape:
        print("%-34s SHAPE MISMATCH got %s ref %s" % (name, got.shape, ref.shape)); return
    err = np.abs(got - ref)
    bad = ~np.isfinite(got)
    print("%-34s max_err %.3e  mean_err %.3e  ref_absmax %.3e  rel %.2e%s" % (
        name, np.nanmax(err), np.nanmean(err), np.abs(ref).max(), np.nanmax(err) / (np.abs(ref).max() + 1e-30),
        "  NONFINITE=%d" % bad.sum() if bad.any() else ""))
    sys.stdout.flush()


# ------------------------------------------------------------------ BDCN
bsd = synth.make_bdcn_state(0)
taps = {}
with torch.no_grad():
    t0 = time.time()
    edge_ref = graph.bdcn_forward(bsd, torch.cat((img, img, img), 1), taps=taps)
    print("oracle bdcn %.1fs" % (time.time() - t0))
edge_model = egn_b200.BDCN()
edge_model.load_state_dict(bsd)
edge_model = edge_model.cuda().eval()
edge_model.micro_batch = 2
x = img.to(dev)
edge = edge_model.edge(x)
torch.cuda.synchronize()
ctx = edge_model.context(dev)
names = ["conv1_1", "conv1_2", "conv2_1", "conv2_2", "conv3_1", "conv3_2", "conv3_3", "conv4_1", "conv4_2", "conv4_3",
         "conv5_1", "conv5_2", "conv5_3"]
for i, n in enumerate(names):
    report("features." + n, ctx.debug_read("features." + n, 2), taps["features"][i].numpy())
nblk = {1: 2, 2: 2, 3: 3, 4: 3, 5: 3}
for st in range(1, 6):
    acc = None
    for j in range(1, nblk[st] + 1):
        d = F.conv2d(taps["msblock%d_%d" % (st, j)], bsd["conv%d_%d_down.weight" % (st, j)])
        acc = d if acc is None else acc + d
    a = F.conv2d(acc, bsd["score_dsn%d.weight" % st]); b = F.conv2d(acc, bsd["score_dsn%d_1.weight" % st])
    report("score%d (no bias)" % st, ctx.debug_read("score%d" % st, 2), torch.cat([a, b], 1).numpy())
report("edge", edge.cpu().numpy(), edge_ref.numpy())
edge3 = edge_model(torch.cat((x, x, x), 1))[-1]
report("edge (3-plane entry)", edge3.cpu().numpy(), edge_ref.numpy())

# ------------------------------------------------------------------ ESF
st = synth.SETTINGS[cfg]
esd = synth.make_esf_state(st, 0)
taps = {}
with torch.no_grad():
    ref = graph.esf_forward(esd, st, img, edge_ref, taps=taps)
model = egn_b200.DenseNet2D(st)
model.load_state_dict(esd)
model = model.cuda().eval()
model.micro_batch = 2
with torch.no_grad():
    op, elPred, latent, loss, elOut = model(x, edge_ref.to(dev), None, None, None, None, None, torch.zeros(2, 4, device=dev), 0, 0)
torch.cuda.synchronize()
ctx = model.context(dev)
nf = 4 if st["add_edge"] else 2
for blk in ["enc.down_block1", "enc.down_block2", "enc.down_block3", "enc.down_block4", "enc.bottleneck"]:
    def cat2(key):
        a = taps[blk + key]
        return torch.cat([a, taps[blk + key + "@edge"]], 0).numpy() if st["add_edge"] else a.numpy()
    if blk == "enc.down_block1":
        h = taps["enc.head"]
        href = torch.cat([h, taps["enc.head@edge"]], 0).numpy() if st["add_edge"] else h.numpy()
        report(blk + ".x (head)", ctx.debug_read(blk + ".x", nf), href)
    report(blk + ".x1", ctx.debug_read(blk + ".x1", nf), cat2(".x1"))
    report(blk + ".x22", ctx.debug_read(blk + ".x22", nf), cat2(".x22"))
    sk = cat2(".skip")
    inter = ctx.debug_read(blk + ".out", nf).shape[1]
    report(blk + ".out", ctx.debug_read(blk + ".out", nf), sk[:, :inter])
    if blk != "enc.bottleneck":
        nxt = {"enc.down_block1": "enc.down_block2", "enc.down_block2": "enc.down_block3",
               "enc.down_block3": "enc.down_block4", "enc.down_block4": "enc.bottleneck"}[blk]
        report(nxt + ".x (TD)", ctx.debug_read(nxt + ".x", nf), cat2(".td"))
    else:
        report("bt (TD)", ctx.debug_read("bt", nf), cat2(".td"))
for i, u in enumerate(["dec.up_block4", "dec.up_block3", "dec.up_block2", "dec.up_block1"]):
    report(u + ".out", ctx.debug_read(u + ".out", 2), taps["dec.up%d" % (4 - i)].numpy())
report("logits", op.cpu().numpy(), ref["op"].numpy())
report("elOut", elOut.cpu().numpy(), ref["elOut"].numpy())
report("elPred", elPred.cpu().numpy(), ref["elPred"].numpy())
report("latent", latent.cpu().numpy(), ref["latent"].numpy())
pred = egn_b200.get_predictions(op, model).numpy()
pref = graph.get_predictions(ref["op"]).numpy()
print("argmax agreement %.5f%%" % (100.0 * (pred == pref).mean()))

if mode == "tc":
    layers = ["features." + n for n in names[1:]]
    for s_ in range(1, 6):
        for j in range(1, nblk[s_] + 1):
            layers += ["msblock%d_%d.conv" % (s_, j), "msblock%d_%d.tail" % (s_, j)]
    ectx = edge_model.context(dev)
    for l in layers:
        try:
            d, r = ectx.conv_selfcheck(l, 2)
            print("selfcheck %-28s max_diff %.3e  ref_absmax %.3e" % (l, d, r))
        except Exception as ex:
            print("selfcheck %-28s FAILED %s" % (l, ex)); break
    layers = ["enc.head.conv2"]
    for blk in ["enc.down_block1", "enc.down_block2", "enc.down_block3", "enc.down_block4", "enc.bottleneck"]:
        layers += [blk + "." + c for c in ("conv1", "conv21", "conv22", "conv31", "conv32", "TD.conv")]
    for u in ["dec.up_block4", "dec.up_block3", "dec.up_block2", "dec.up_block1"]:
        layers += [u + "." + c for c in ("pre", "conv11", "conv12", "conv21", "conv22")]
    layers += ["dec.final.conv1", "dec.final.conv2", "elReg.c1"]
    for l in layers:
        try:
            d, r = ctx.conv_selfcheck(l, 2)
            print("selfcheck %-28s max_diff %.3e  ref_absmax %.3e" % (l, d, r))
        except Exception as ex:
            print("selfcheck %-28s FAILED %s" % (l, ex)); break
print("DONE", mode, cfg)
