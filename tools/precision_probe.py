"""CPU emulation of operand-precision policies for the tensor-core convolutions.

Test/analysis tooling (uses the oracle): every conv of oracle/graph.py is re-run with its input
and weight operands rounded the way a given storage/MMA policy would round them (fp32 accumulate),
and the outputs are compared with the fp32 oracle on the bars of BASELINE.json
(argmax agreement >= 99.9 %, centres within 0.25 px, ellipse parameters within 1e-2 relative).

    python tools/precision_probe.py [frames]

Policies are "<A>/<W>" per net: single = one fp16 (h) or bf16 (b) plane, x2 = hi+lo split.
MMA cost per policy: A1*W1 = 1, A2*W1 = 2, A2*W2 = 3 (hi*hi + lo*hi + hi*lo).
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import graph, synth  # noqa: E402


def rnd(x, kind):
    if kind == "f32":
        return x
    if kind == "b":
        return x.bfloat16().float()
    if kind == "h":
        return x.half().float()
    if kind == "b2":
        h = x.bfloat16().float()
        return h + (x - h).bfloat16().float()
    if kind == "h2":
        h = x.half().float()
        return h + (x - h).half().float()
    if kind == "hb":            # fp16 hi + bf16 lo
        h = x.half().float()
        return h + (x - h).bfloat16().float()
    raise ValueError(kind)


class FShim:
    """Stands in for torch.nn.functional inside oracle.graph with operand-rounded conv2d."""

    def __init__(self):
        self.a, self.w = "f32", "f32"

    def __getattr__(self, name):
        return getattr(F, name)

    def conv2d(self, x, w, b=None, **kw):
        return F.conv2d(rnd(x, self.a), rnd(w, self.w), b, **kw)


def run(policy_bdcn, policy_esf, bsd, esd, st, img, shim):
    with torch.no_grad():
        shim.a, shim.w = policy_bdcn
        edge = graph.calc_edge(bsd, img)
        shim.a, shim.w = policy_esf
        out = graph.esf_forward(esd, st, img, edge)
    shim.a = shim.w = "f32"
    return edge, out


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    torch.set_num_threads(os.cpu_count() or 1)
    st = synth.SETTINGS["baseline_edge"]
    bsd, esd = synth.make_bdcn_state(0), synth.make_esf_state(st, 0)
    fr = np.load(os.path.join(ROOT, "tests", "golden", "frames_u8.npz"))["frames"][:n]
    real = torch.from_numpy(np.stack([graph.preprocess_frame_u8(f) for f in fr]))[:, None]
    eyes = torch.from_numpy(synth.synthetic_eye_batch(0, n)["img"])
    img = torch.cat([real, eyes, synth.randn_frames(n, seed=7)], 0)
    shim = FShim()
    graph.F = shim
    e0, r0 = run(("f32", "f32"), ("f32", "f32"), bsd, esd, st, img, shim)
    p0 = graph.get_predictions(r0["op"])
    top2 = r0["op"].topk(2, dim=1)[0]
    print("frames %d; pixels with top-2 margin < 0.05: %.2f%%" % (img.shape[0], 100 * ((top2[:, 0] - top2[:, 1]) < 0.05).float().mean()))
    pols = [
        (("b2", "b2"), ("b2", "b2")),
        (("b", "b"), ("b", "b")),
        (("h", "h"), ("h", "h")),
        (("h", "h"), ("b2", "b2")),
        (("b", "b"), ("b2", "b2")),
        (("h", "h"), ("h2", "h")),
        (("h", "h"), ("h2", "h2")),
        (("h2", "h"), ("h2", "h")),
        (("h2", "h"), ("b2", "b2")),
        (("h", "h2"), ("b2", "b2")),
        (("h", "h"), ("hb", "h")),
    ]
    print("%-22s %-9s %-9s %-10s %-10s %-10s" % ("BDCN a/w | ESF a/w", "edge_err", "argmax%", "centre_px", "ell_rel", "min agree"))
    for pb, pe in pols:
        e, r = run(pb, pe, bsd, esd, st, img, shim)
        p = graph.get_predictions(r["op"])
        per = (p == p0).float().flatten(1).mean(1)
        cpx = ((r["elPred"] - r0["elPred"])[:, [0, 1, 5, 6]].abs() * torch.tensor([160., 120., 160., 120.])).max().item()
        par = [2, 3, 4, 7, 8, 9]
        rel = ((r["elOut"] - r0["elOut"])[:, par].abs() / r0["elOut"][:, par].abs().clamp_min(1e-2)).max().item()
        print("%-22s %-9.2e %-9.4f %-10.4f %-10.2e %-10.4f" % ("%s/%s | %s/%s" % (pb + pe), (e - e0).abs().max().item(),
                                                                100 * per.mean().item(), cpx, rel, 100 * per.min().item()))


if __name__ == "__main__":
    main()
