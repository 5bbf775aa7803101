"""Generates tests/golden/loss.npz with the REAL reference (build container only; needs /root/reference):
get_allLoss (models/RITnet_v2.py:372-440) and its three per-sample segmentation losses (loss.py) on
seeded inputs, and asserts oracle/graph.py's restatement against them on the same tensors.

    python oracle/make_golden_loss.py

Inputs (all reproducible from oracle/synth.py, so the fixture only stores the reference's outputs):
the labels / centres / ellipses of synthetic eyes 500..503, cond[:,1] = [0,0,1,0] (sample 2 has no
mask), logits = synth.smooth_logits(label, 5), spatial weights / distance maps = synth.loss_maps(label, 6)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import graph, ref_harness, synth      # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    ref_harness.install_shims()
    from models.RITnet_v2 import get_allLoss
    import loss as ref_loss
    eb = synth.synthetic_eye_batch(500, 4)
    cond = torch.from_numpy(eb["cond"]).clone().float()
    cond[2, 1] = 1
    op = synth.smooth_logits(eb["label"], seed=5)
    el_out = torch.from_numpy(np.random.RandomState(9).uniform(-0.5, 0.5, (4, 10)).astype(np.float32))
    tgt = torch.from_numpy(eb["label"]).long()
    pc = torch.from_numpy(eb["pupil_center"]).float()
    en = torch.from_numpy(eb["elNorm"]).float()
    sw, dm = synth.loss_maps(eb["label"], seed=6)
    out = {}
    for alpha in (0.0, 0.5, 1.0):
        total, pcs = get_allLoss(op, el_out, tgt, pc, en, sw, dm, cond, torch.zeros(4, dtype=torch.long), alpha)
        mine, pcs2 = graph.all_loss(op, el_out, tgt, pc, en, sw, dm, cond, alpha)
        assert abs(float(total) - float(mine)) <= 1e-5 * max(1.0, abs(float(total))), (alpha, float(total), float(mine))
        assert torch.allclose(pcs, pcs2, atol=1e-6)
        out["total_a%d" % int(alpha * 10)] = np.float64(float(total))
    cond_none = cond.clone(); cond_none[:, 1] = 1
    total, _ = get_allLoss(op, el_out, tgt, pc, en, sw, dm, cond_none, torch.zeros(4, dtype=torch.long), 0.5)
    mine, _ = graph.all_loss(op, el_out, tgt, pc, en, sw, dm, cond_none, 0.5)
    assert abs(float(total) - float(mine)) <= 1e-5 * max(1.0, abs(float(total)))
    out["total_nomask"] = np.float64(float(total))
    sl = [float(ref_loss.SurfaceLoss(op[i:i + 1], dm[i:i + 1])) for i in range(4)]
    ce = [float(ref_loss.wCE(op[i], tgt[i], sw[i])) for i in range(4)]
    gd = [float(ref_loss.GDiceLoss(op[i:i + 1], tgt[i:i + 1], torch.nn.functional.softmax)) for i in range(4)]
    for i in range(4):
        assert abs(sl[i] - float(graph.surface_loss(op[i], dm[i]))) < 1e-6
        assert abs(ce[i] - float(graph.wce_loss(op[i], tgt[i], sw[i]))) < 1e-6
        assert abs(gd[i] - float(graph.gdice_loss(op[i], tgt[i]))) < 1e-6
    np.savez_compressed(os.path.join(OUT, "loss.npz"), surface=np.array(sl), wce=np.array(ce), gdice=np.array(gd),
                        el_out=el_out.numpy(), **out)
    print("loss golden", out, sl, ce, gd)


if __name__ == "__main__":
    main()
