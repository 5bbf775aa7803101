"""Generates tests/golden/* by running the UNMODIFIED reference (build container only).

    python -m oracle.make_golden

Needs /root/reference.  The fixtures are committed; this script is the recipe.
"""
import json
import os
import sys
import types
import contextlib
import io

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth, ref_harness  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
CONFIGS = ["baseline", "baseline_edge", "baseline_adain", "baseline_adain_edge",
           "baseline_input_concat", "baseline_only_edge"]


def real_frames():
    """evaluate.py:243-249: split 240x640 BGR frame into two eyes, BGR->gray."""
    import cv2
    cap = cv2.VideoCapture(os.path.join(ref_harness.REF, "videos", "example1.avi"))
    frames = []
    idx = 0
    want = {0, 150, 300, 449}
    while True:
        ret, fr = cap.read()
        if not ret:
            break
        if idx in want:
            for i in range(2):
                frames.append(cv2.cvtColor(fr[:, i * 320:(i + 1) * 320, :], cv2.COLOR_BGR2GRAY))
        idx += 1
    cap.release()
    return np.stack(frames).astype(np.uint8)


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    BDCN, DenseNet2D, ref_utils, ref_hf = ref_harness.load_reference_modules()

    # 1. key/shape listing of the real modules
    keys = {}
    with contextlib.redirect_stdout(io.StringIO()):
        keys["bdcn"] = {k: list(v.shape) for k, v in BDCN().state_dict().items()}
        for c in CONFIGS:
            keys[c] = {k: list(v.shape) for k, v in DenseNet2D(dict(synth.SETTINGS[c])).state_dict().items()}
    with open(os.path.join(OUT, "state_keys.json"), "w") as f:
        json.dump(keys, f, indent=0, sort_keys=True)

    # 2. real frames
    fr = real_frames()
    np.savez_compressed(os.path.join(OUT, "frames_u8.npz"), frames=fr)
    print("frames", fr.shape)

    # 3. forward goldens (B=2: one real crop + one synthetic eye)
    eye = synth.synthetic_eye(7)
    z0 = (fr[2].astype(np.float64) - fr[2].mean()) / fr[2].std()
    img = torch.from_numpy(np.stack([z0.astype(np.float32)[None], eye["img"]]))
    bsd = synth.make_bdcn_state(0)
    np.savez_compressed(os.path.join(OUT, "fwd_input.npz"), img=img.numpy())
    for c in CONFIGS:
        st = synth.SETTINGS[c]
        esd = synth.make_esf_state(st, 0)
        r = ref_harness.run_reference(st, bsd, esd, img)
        cond_nomask = torch.zeros((2, 4)); cond_nomask[:, 1] = 1
        r2 = ref_harness.run_reference(st, bsd, esd, img, cond=cond_nomask)
        op = r["op"].numpy()
        srt = np.sort(op, 1)
        margin = srt[:, -1] - srt[:, -2]
        print(c, "edge mean %.4f std %.4f | logits absmax %.2f margin<0.05: %.3f%% | elOut %s" % (
            r["edge"].mean(), r["edge"].std(), np.abs(op).max(), 100 * (margin < 0.05).mean(),
            np.round(r["elOut"][0].numpy(), 3)))
        np.savez_compressed(
            os.path.join(OUT, f"fwd_{c}.npz"),
            edge=r["edge"].numpy() if c == "baseline_edge" else r["edge"].numpy()[:, :, ::4, ::4],
            op_s4=op[:, :, ::4, ::4], pred=r["pred"].numpy().astype(np.uint8),
            margin_f16=margin.astype(np.float16),
            elPred=r["elPred"].numpy(), elOut=r["elOut"].numpy(), latent=r["latent"].numpy(),
            elPred_nomask=r2["elPred"].numpy(),
            op_mean=op.mean((2, 3)), op_absmean=np.abs(op).mean((2, 3)))

    # 4. ellipse geometry goldens
    rng = np.random.RandomState(3)
    Hm = np.array([[320 / 2, 0, 320 / 2], [0, 240 / 2, 240 / 2], [0, 0, 1.0]])
    params, outs = [], []
    for i in range(16):
        p = np.array([rng.uniform(-0.5, 0.5), rng.uniform(-0.5, 0.5), rng.uniform(0.1, 0.6),
                      rng.uniform(0.1, 0.6), rng.uniform(-1.5, 1.5)])
        params.append(p)
        outs.append(ref_hf.my_ellipse(p).transform(Hm)[0])
    masks, inits, refined, ious = [], [], [], []
    for i in range(6):
        e = synth.synthetic_eye(20 + i)
        for cls, row in ((1, 0), (2, 1)):
            m = (e["label"] >= cls) if cls == 1 else (e["label"] == 2)
            if cls == 1:
                m = e["label"] == 1
            px = ref_hf.my_ellipse(e["elNorm"][row].astype(np.float64)).transform(Hm)[0][:-1]
            px = px * np.array([1.0, 1.0, rng.uniform(0.85, 1.15), rng.uniform(0.85, 1.15), 1.0])
            px[4] += rng.uniform(-0.2, 0.2)
            px[0] += rng.uniform(-2, 2)
            seg = torch.from_numpy(m)
            mesh = ref_utils.create_meshgrid(240, 320, normalized_coordinates=True)
            init_deg = np.array([px[0], px[1], px[2], px[3], px[4] * 180. / 3.14159])
            ious.append(ref_utils.calc_ell_iou(seg, init_deg.copy(), mesh, False, True))
            out = ref_utils.search_proper_parameter_iou_for_our_data(seg, px.copy())
            masks.append(np.packbits(m)); inits.append(px); refined.append(out)
    np.savez_compressed(os.path.join(OUT, "ellipse.npz"), params=np.array(params), transformed=np.array(outs),
                        masks=np.array(masks), inits=np.array(inits), refined=np.array(refined),
                        init_iou=np.array(ious))
    print("ellipse refine ok", np.array(refined)[0], np.array(inits)[0])

    # 5. metric goldens
    eb = synth.synthetic_eye_batch(40, 6)
    lab = eb["label"].copy()
    pred = lab.copy()
    nz = rng.rand(*pred.shape) < 0.05
    pred[nz] = rng.randint(0, 3, nz.sum())
    lab[4][lab[4] == 2] = 1            # sample without pupil class -> NaN slot
    cond = np.array([0, 0, 1, 0, 0, 0], np.float32)
    miou, per, by = ref_utils.getSeg_metrics(lab, pred, cond)
    ptrue = eb["pupil_center"]; ppred = eb["elNorm"][:, 1, :2] + rng.normal(0, 0.02, (6, 2)).astype(np.float32)
    pd, pds = ref_utils.getPoint_metric(ptrue, ppred, cond, (240, 320), True)
    np.savez_compressed(os.path.join(OUT, "metrics.npz"), label=lab.astype(np.uint8), pred=pred.astype(np.uint8),
                        cond=cond, miou=miou, per=per, by=by, ptrue=ptrue, ppred=ppred, pd=pd, pds=pds)
    print("metrics", miou, per, pd)
    tot = sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT))
    print("golden bytes", tot)


if __name__ == "__main__":
    main()
