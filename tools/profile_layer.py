"""Runs the ESF-Net (baseline_edge) forward twice on a small batch so that ncu can capture single
conv_tc_kernel launches of the second (warm) pass:

    ncu --set full --clock-control none --import-source on -k regex:conv_tc -s <54 + i> -c <n> \
        -o gpurun_out/prof python tools/profile_layer.py [esf|bdcn] [batch]

conv_tc launch order, ESF: 0 enc.head.conv2, then per encoder block conv1, conv21, conv22, conv31,
conv32, TD.conv (1..30), then per up block pre, conv11, conv12, conv21, conv22 (31..50), 51 dec.final.conv1,
52 dec.final.conv2, 53 elReg.c1.
BDCN: per VGG layer i: [features.conv(i) for i>0], msblock.conv, msblock.tail (38 launches)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import egn_b200
from oracle import synth

net = sys.argv[1] if len(sys.argv) > 1 else "esf"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dev = torch.device("cuda:0")
x = synth.randn_frames(B, seed=3).to(dev)
if net == "bdcn":
    m = egn_b200.BDCN(); m.load_state_dict(synth.make_bdcn_state(0)); m = m.cuda().eval(); m.micro_batch = B
    for _ in range(2):
        m.edge(x)
else:
    st = synth.SETTINGS["baseline_edge"]
    m = egn_b200.DenseNet2D(st); m.load_state_dict(synth.make_esf_state(st, 0)); m = m.cuda().eval(); m.micro_batch = B
    e = torch.rand(B, 1, 240, 320, device=dev)
    for _ in range(2):
        m.infer(x, e, None)
torch.cuda.synchronize()
print("done")
