"""Probe: does running two half-batches concurrently on two streams (each engine's persistent grids capped at
half the SMs, EGN_NUM_SMS) beat one full batch?  HBM-bound layers of one lane could overlap tensor-bound layers
of the other; the step is power-capped, so the answer has to be measured.

    EGN_NUM_SMS=74 python tools/colane_probe.py 2 128     # two lanes of 128 frames
    python tools/colane_probe.py 1 256                     # one lane of 256 frames
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import egn_b200
from oracle import synth

lanes = int(sys.argv[1]); B = int(sys.argv[2]); steps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
dev = torch.device("cuda:0")
st = synth.SETTINGS["baseline_edge"]
bsd = synth.make_bdcn_state(0); esd = synth.make_esf_state(st, 0)
models = []
for l in range(lanes):
    em = egn_b200.BDCN(); em.load_state_dict(bsd); em = em.cuda().eval(); em.micro_batch = B
    m = egn_b200.DenseNet2D(st); m.load_state_dict(esd); m = m.cuda().eval(); m.micro_batch = B
    models.append((em, m, torch.cuda.Stream(), synth.randn_frames(B, seed=3 + l).to(dev)))


def step():
    for em, m, s, x in models:
        s.wait_stream(torch.cuda.current_stream())
    for em, m, s, x in models:
        with torch.cuda.stream(s):
            e = em.edge(x)
    # second pass enqueues the ESF-Nets so that lane 0's ESF-Net can overlap lane 1's BDCN
    outs = []
    for em, m, s, x in models:
        with torch.cuda.stream(s):
            e = em.edge(x) if False else e
    return outs


def step2(offset):
    edges = []
    for i, (em, m, s, x) in enumerate(models):
        s.wait_stream(torch.cuda.current_stream())
    if offset:
        # lane 0: BDCN, ESF; lane 1: ESF(prev edge), BDCN  -> the two nets are in anti-phase
        em0, m0, s0, x0 = models[0]; em1, m1, s1, x1 = models[1]
        with torch.cuda.stream(s0):
            e0 = em0.edge(x0)
        with torch.cuda.stream(s1):
            m1.infer(x1, step2.prev_e1, None)
        with torch.cuda.stream(s0):
            m0.infer(x0, e0, None)
        with torch.cuda.stream(s1):
            step2.prev_e1 = em1.edge(x1)
    else:
        for em, m, s, x in models:
            with torch.cuda.stream(s):
                e = em.edge(x)
                m.infer(x, e, None)
    for em, m, s, x in models:
        torch.cuda.current_stream().wait_stream(s)


for offset in ([0, 1] if lanes == 2 else [0]):
    if offset:
        with torch.cuda.stream(models[1][2]):
            step2.prev_e1 = models[1][0].edge(models[1][3])
        torch.cuda.synchronize()
    for _ in range(2):
        step2(offset)
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(steps):
        step2(offset)
    t1.record(); torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / steps
    print("lanes %d x %d frames, anti-phase %d, EGN_NUM_SMS=%s: %.2f ms/step, %.1f frames/s" % (
        lanes, B, offset, os.environ.get("EGN_NUM_SMS", "-"), ms, lanes * B / ms * 1e3))
