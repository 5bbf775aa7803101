"""Operand-precision policies measured on the GPU kernels themselves (the CPU emulation of
tools/precision_probe_layers.py was the round-1 evidence).

Every policy lowers a set of layers from three tensor-core products per MAC (hi*hi + lo*hi + hi*lo) to two through
the engine's EGN_PRODUCTS probe knob (mode 1: activations act as bf16, mode 2: weights act as bf16) and is compared
with the fp32 CPU oracle on BASELINE.json's bars over real eye crops, synthetic eyes and noise frames.  A policy
would be adopted only if ALL bars hold with a 2x margin (argmax >= 99.95 %, centres <= 0.125 px, ellipse <= 5e-3).

    python tools/gpu_precision_probe.py [frames_per_kind] > profiles/r02_gpu_precision_probe.txt

Each policy runs in its own process (the knob is read when the context is created).  Test infrastructure: imports oracle/."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

POLICIES = [
    ("all three products (shipping policy)", ""),
    ("VGG stage 5 (conv5_x): W as bf16", "features.conv5=2"),
    ("VGG stage 5 (conv5_x): A as bf16", "features.conv5=1"),
    ("VGG stages 4-5: W as bf16", "features.conv4=2,features.conv5=2"),
    ("VGG stages 4-5: A as bf16", "features.conv4=1,features.conv5=1"),
    ("MSBlock .conv stages 4-5: W as bf16", "msblock4=2,msblock5=2"),
    ("MSBlock tails stages 1-2 (phase lattice): A as bf16", "msblock1_1.tail=1,msblock1_2.tail=1,msblock2_1.tail=1,msblock2_2.tail=1"),
    ("all of BDCN: W as bf16 (tails: A as bf16)", "msblock1_1.tail=1,msblock1_2.tail=1,msblock2_1.tail=1,msblock2_2.tail=1,features=2,msblock=2"),
    ("ESF-Net decoder: W as bf16", "dec.=2"),
    ("ESF-Net encoder blocks 3-4 + bottleneck: W as bf16", "enc.down_block3=2,enc.down_block4=2,enc.bottleneck=2"),
    ("ESF-Net style / head only: W as bf16", "elReg=2"),
]


def worker(n):
    import numpy as np
    import torch
    import egn_b200
    from oracle import graph, synth
    dev = torch.device("cuda:0")
    st = synth.SETTINGS["baseline_edge"]
    bsd, esd = synth.make_bdcn_state(0), synth.make_esf_state(st, 0)
    fr = np.load(os.path.join(ROOT, "tests", "golden", "frames_u8.npz"))["frames"]
    real = torch.from_numpy(np.stack([graph.preprocess_frame_u8(f) for f in fr]))[:, None]
    reps = (n + len(fr) - 1) // len(fr)
    # more real-frame variety than the 8 committed crops: horizontal / vertical flips and 90-pixel shifts of them
    aug = [real, real.flip(3), real.flip(2), real.roll(90, 3)]
    real = torch.cat(aug[:max(1, min(4, reps))], 0)[:n]
    eyes = torch.from_numpy(synth.synthetic_eye_batch(0, n)["img"])
    img = torch.cat([real, eyes, synth.randn_frames(n, seed=7)], 0).float()
    em = egn_b200.BDCN(); em.load_state_dict(bsd); em = em.cuda().eval(); em.micro_batch = 16
    m = egn_b200.DenseNet2D(st); m.load_state_dict(esd); m = m.cuda().eval(); m.micro_batch = 16
    with torch.no_grad():
        e = em.edge(img.to(dev))
        op, elPred, latent, loss, elOut = m(img.to(dev), e, None, None, None, None, None, torch.zeros(len(img), 4, device=dev), 0, 0)
        pred = egn_b200.get_predictions(op, m)
    torch.cuda.synchronize()
    cache = os.path.join(ROOT, "gpurun_out", "probe_oracle_%d.npz" % n)
    if os.path.exists(cache):
        z = np.load(cache); e0, p0, ep0, eo0 = (torch.from_numpy(z[k]) for k in ("e", "p", "ep", "eo"))
    else:
        torch.set_num_threads(os.cpu_count() or 1)
        with torch.no_grad():
            e0 = graph.calc_edge(bsd, img)
            r0 = graph.esf_forward(esd, st, img, e0)
        p0, ep0, eo0 = graph.get_predictions(r0["op"]), r0["elPred"], r0["elOut"]
        os.makedirs(os.path.dirname(cache), exist_ok=True)
        np.savez(cache, e=e0.numpy(), p=p0.numpy(), ep=ep0.numpy(), eo=eo0.numpy())
    per = (pred.cpu() == p0).float().flatten(1).mean(1)
    cpx = ((elPred.cpu() - ep0)[:, [0, 1, 5, 6]].abs() * torch.tensor([160., 120., 160., 120.])).max().item()
    par = [2, 3, 4, 7, 8, 9]
    rel = ((elOut.cpu() - eo0)[:, par].abs() / eo0[:, par].abs().clamp_min(1e-2)).max().item()
    info = m.context(dev).info(); einfo = em.context(dev).info()
    print(json.dumps({"edge_err": (e.cpu() - e0).abs().max().item(), "argmax": 100 * per.mean().item(), "min_agree": 100 * per.min().item(),
                      "centre_px": cpx, "ell_rel": rel, "frames": len(img), "lowered": info["lowered_layers"] + einfo["lowered_layers"]}))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    print("# GPU kernels, %d real crops (8 committed crops + flips / shifts) + %d synthetic eyes + %d noise frames; baseline_edge, synthetic checkpoints" % (n, n, n))
    print("# bars: argmax >= 99.9 %, centres <= 0.25 px, ellipse <= 1e-2 relative (floor 1e-2); adoption needs a 2x margin on every bar")
    print("%-54s %-8s %-9s %-9s %-10s %-10s %-9s %s" % ("policy", "layers", "edge_err", "argmax%", "min agree", "centre_px", "ell_rel", "verdict"))
    for name, pol in POLICIES:
        env = dict(os.environ)
        env.pop("EGN_PRODUCTS", None)
        if pol:
            env["EGN_PRODUCTS"] = pol
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--worker", str(n)], env=env, capture_output=True, text=True)
        line = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
        if not line:
            print("%-54s FAILED: %s" % (name, (out.stderr or out.stdout)[-300:].replace("\n", " | ")))
            continue
        r = json.loads(line[-1])
        ok = r["min_agree"] >= 99.9 and r["centre_px"] <= 0.25 and r["ell_rel"] <= 1e-2
        ok2 = r["min_agree"] >= 99.95 and r["centre_px"] <= 0.125 and r["ell_rel"] <= 5e-3
        print("%-54s %-8d %-9.2e %-9.4f %-10.4f %-10.4f %-9.2e %s" % (name, r["lowered"], r["edge_err"], r["argmax"], r["min_agree"], r["centre_px"],
                                                                     r["ell_rel"], "meets bars with 2x margin" if ok2 else ("meets bars, no margin" if ok else "FAILS")))
        sys.stdout.flush()


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--worker":
        worker(int(sys.argv[2]))
    else:
        main()
