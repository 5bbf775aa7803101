"""Generates tests/golden/planted_<config>.npz: the reference's own outputs for the frames that the
benchmarked-configuration pin plants into bench.py's 256-frame batch (VERDICT r1 "next" #1).

TEST INFRASTRUCTURE.  Runs in the build container only (imports the unmodified reference from
/root/reference through oracle/ref_harness.py).  Planted frames: the two frames of fwd_input.npz (one
real crop + one synthetic eye) followed by the first eight real crops of frames_u8.npz, z-scored like
evaluate.py:102-103.  Weights: oracle/synth.py seed 0 (the ones bench.py uses).

    python -m oracle.make_golden_planted
"""
import os

import numpy as np
import torch

from . import ref_harness, synth

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
CONFIGS = ["baseline_edge", "baseline_adain_edge"]


def planted_frames():
    img = np.load(os.path.join(OUT, "fwd_input.npz"))["img"]
    fr = np.load(os.path.join(OUT, "frames_u8.npz"))["frames"][:8].astype(np.float64)
    z = np.stack([((f - f.mean()) / f.std()).astype(np.float32)[None] for f in fr])
    return np.concatenate([img, z], 0)


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    x = torch.from_numpy(planted_frames())
    bsd = synth.make_bdcn_state(0)
    for c in CONFIGS:
        st = synth.SETTINGS[c]
        r = ref_harness.run_reference(st, bsd, synth.make_esf_state(st, 0), x)
        np.savez_compressed(os.path.join(OUT, "planted_%s.npz" % c), pred=r["pred"].numpy().astype(np.uint8),
                            elPred=r["elPred"].numpy(), elOut=r["elOut"].numpy(), latent=r["latent"].numpy(),
                            edge_s4=r["edge"].numpy()[:, :, ::4, ::4])
        print(c, "pred classes", np.bincount(r["pred"].numpy().ravel(), minlength=3), "elOut[0]", np.round(r["elOut"][0].numpy(), 3))


if __name__ == "__main__":
    main()
