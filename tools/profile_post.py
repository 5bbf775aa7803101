"""Runs the bandwidth-bound post-processing kernels (K6 / K7 / loss slot / u8 ingest) twice on a
256-frame batch so that ncu can capture the warm launches:

    ncu --set full --clock-control none -k regex:"seg_post_kernel|seg_counts_kernel|seg_loss_kernel|preprocess_u8_kernel" \
        -s 4 -c 4 -o gpurun_out/post python tools/profile_post.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import egn_b200

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda:0")
ctx = egn_b200.Context(dev, None, 1)
g = torch.Generator(device="cuda").manual_seed(0)
logits = torch.randn(B, 3, 240, 320, device=dev, generator=g)
labels = torch.randint(0, 3, (B, 240, 320), device=dev, dtype=torch.uint8, generator=g)
frames = torch.randint(0, 256, (B, 240, 320), device=dev, dtype=torch.uint8, generator=g)
sw = torch.rand(B, 240, 320, device=dev, generator=g) + 1
dm = torch.randn(B, 3, 240, 320, device=dev, generator=g)
cond = torch.zeros(B, 4, device=dev)
el = torch.zeros(B, 10, device=dev)
pc = torch.full((B, 2), 100.0, device=dev)
en = torch.zeros(B, 2, 5, device=dev)
acc = egn_b200.MetricAccumulator(dev)
for _ in range(2):
    x = ctx.preprocess_u8(frames)
    am, ep = ctx.seg_post(logits, el, cond)
    ctx.metrics_accumulate(am, labels, cond, acc.acc, pc, pc, el, ep)
    loss = ctx.forward_loss(logits, labels, sw, dm, cond, pc, en, el, ep, 0.5)
torch.cuda.synchronize()
print("done", float(loss.item()))
