"""Per-layer operand-precision probe for the BDCN edge net (CPU emulation, test/analysis tooling).

The deep VGG stages (conv4_x / conv5_x, 512 channels at 30x40) only reach the edge map through the
k16/s8 transposed-convolution upsamplers, which average their rounding noise over 16x16 windows, so
they may tolerate single-plane operands while everything else keeps the three-product split.  Every
conv of oracle/graph.py is re-run with operands rounded per a (layer -> policy) rule and the result is
compared with the fp32 oracle on BASELINE.json's bars.

    python tools/precision_probe_layers.py [frames]
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from oracle import graph, synth  # noqa: E402
from precision_probe import rnd  # noqa: E402


class LayerShim:
    def __init__(self):
        self.rule = None            # callable(x, w, kw) -> (a_kind, w_kind)

    def __getattr__(self, name):
        return getattr(F, name)

    def conv2d(self, x, w, b=None, **kw):
        a, wk = self.rule(x, w, kw) if self.rule else ("f32", "f32")
        return F.conv2d(rnd(x, a), rnd(w, wk), b, **kw)


def bdcn_rule(deep, stage_min):
    """deep: policy for VGG convs (and optionally MSBlock convs) whose input map is <= 30 rows (stage >= 4)
    or <= 60 rows (stage >= 3); everything else split-bf16."""
    def rule(x, w, kw):
        h = x.shape[2]
        lim = {3: 60, 4: 30, 5: 29}[stage_min]
        is_vgg = w.shape[0] >= 256 and w.shape[2] == 3
        if is_vgg and h <= lim and (stage_min < 5 or kw.get("dilation", 1) in (2, (2, 2))):
            return deep
        return ("b2", "b2")
    return rule


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    torch.set_num_threads(os.cpu_count() or 1)
    st = synth.SETTINGS["baseline_edge"]
    bsd, esd = synth.make_bdcn_state(0), synth.make_esf_state(st, 0)
    fr = np.load(os.path.join(ROOT, "tests", "golden", "frames_u8.npz"))["frames"][:n]
    real = torch.from_numpy(np.stack([graph.preprocess_frame_u8(f) for f in fr]))[:, None]
    eyes = torch.from_numpy(synth.synthetic_eye_batch(0, n)["img"])
    img = torch.cat([real, eyes, synth.randn_frames(n, seed=7)], 0)
    shim = LayerShim()
    graph.F = shim
    with torch.no_grad():
        e0 = graph.calc_edge(bsd, img)
        r0 = graph.esf_forward(esd, st, img, e0)
    p0 = graph.get_predictions(r0["op"])
    split = lambda x, w, kw: ("b2", "b2")
    cases = [("all split-bf16 (shipping policy)", split),
             ("VGG stage>=4 single bf16", bdcn_rule(("b", "b"), 4)),
             ("VGG stage>=4 A split / W bf16", bdcn_rule(("b2", "b"), 4)),
             ("VGG stage>=4 A bf16 / W split", bdcn_rule(("b", "b2"), 4)),
             ("VGG stage>=4 single fp16", bdcn_rule(("h", "h"), 4)),
             ("VGG stage 5 single bf16", bdcn_rule(("b", "b"), 5)),
             ("VGG stage>=3 single bf16", bdcn_rule(("b", "b"), 3)),
             ("VGG stage>=3 single fp16", bdcn_rule(("h", "h"), 3)),
             ("VGG stage>=4 A fp16x2 / W fp16", bdcn_rule(("h2", "h"), 4)),
             ("VGG stage>=4 A fp16 / W fp16x2", bdcn_rule(("h", "h2"), 4)),
             ("VGG stage 5 single fp16", bdcn_rule(("h", "h"), 5)),
             ("VGG stage>=3 A fp16x2 / W fp16", bdcn_rule(("h2", "h"), 3))]
    print("%-36s %-9s %-9s %-10s %-10s %-10s" % ("policy", "edge_err", "argmax%", "centre_px", "ell_rel", "min agree"))
    for name, rule in cases:
        with torch.no_grad():
            shim.rule = rule
            e = graph.calc_edge(bsd, img)
            shim.rule = split
            r = graph.esf_forward(esd, st, img, e)
        p = graph.get_predictions(r["op"])
        per = (p == p0).float().flatten(1).mean(1)
        cpx = ((r["elPred"] - r0["elPred"])[:, [0, 1, 5, 6]].abs() * torch.tensor([160., 120., 160., 120.])).max().item()
        par = [2, 3, 4, 7, 8, 9]
        rel = ((r["elOut"] - r0["elOut"])[:, par].abs() / r0["elOut"][:, par].abs().clamp_min(1e-2)).max().item()
        print("%-36s %-9.2e %-9.4f %-10.4f %-10.2e %-10.4f" % (name, (e - e0).abs().max().item(), 100 * per.mean().item(), cpx, rel,
                                                               100 * per.min().item()))
        sys.stdout.flush()


if __name__ == "__main__":
    main()
