import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch, egn_b200
from oracle import synth
dev = torch.device("cuda", 0)
st = synth.SETTINGS["baseline_edge"]
em = egn_b200.BDCN(); em.load_state_dict(synth.make_bdcn_state(0)); em = em.cuda().eval(); em.micro_batch = int(sys.argv[1]) if len(sys.argv) > 1 else 1
m = egn_b200.DenseNet2D(st); m.load_state_dict(synth.make_esf_state(st, 0)); m = m.cuda().eval(); m.micro_batch = int(sys.argv[1]) if len(sys.argv) > 1 else 1
fr = np.load("tests/golden/frames_u8.npz")["frames"]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
f = torch.from_numpy(np.stack([fr[i % len(fr)] for i in range(B)])).to(dev)
for _ in range(3):
    x = egn_b200.preprocess_frames_u8(f, dev)
    e = em.edge(x)
    lg, eo, lat, am, ep = m.infer(x, e, None)
    ell = m.context(dev).ellipse_refine(am, ep, True)
torch.cuda.synchronize()
