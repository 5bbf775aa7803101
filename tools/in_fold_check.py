"""Algebra check for a next-round fusion (CPU, float64): the InstanceNorm in front of each dense block's
conv1 (models/RITnet_v2.py:60, `x1 = lrelu(conv1(IN(x)))`) can be folded into per-frame weights and a
position-dependent bias, which would remove the `instnorm_apply_kernel` pass on x (15 us/frame today):

    conv1(IN(x))[n, co, p] = sum_{t in taps, in-bounds at p} sum_c (W[co,c,t] * s[n,c]) * x[n,c,p+t]
                             + b[co] - sum_{t in-bounds at p} T[n,co,t],      T[n,co,t] = sum_c W[co,c,t] * mu[n,c] * s[n,c]

with s = 1/sqrt(var + eps), mu the per-(frame, channel) mean.  Zero padding lives in IN(x)-space, so the
bias correction only sums the taps that are inside the image at p (9 border classes for a 3x3).

    python tools/in_fold_check.py
"""
import torch
import torch.nn.functional as F

torch.manual_seed(0)
N, C, Co, H, W = 2, 6, 5, 9, 11
x = torch.randn(N, C, H, W, dtype=torch.float64) * 3 + 1.5
w = torch.randn(Co, C, 3, 3, dtype=torch.float64)
b = torch.randn(Co, dtype=torch.float64)
ref = F.conv2d(F.instance_norm(x, eps=1e-5), w, b, padding=1)

mu = x.mean((2, 3))
s = 1.0 / torch.sqrt(x.var((2, 3), unbiased=False) + 1e-5)
out = torch.empty_like(ref)
ones = torch.ones(1, 1, H, W, dtype=torch.float64)
for n in range(N):
    wn = w * s[n].view(1, C, 1, 1)                                   # per-frame scaled weights
    y = F.conv2d(x[n:n + 1], wn, None, padding=1)
    T = (w * (mu[n] * s[n]).view(1, C, 1, 1)).sum(1)                 # [Co,3,3]
    inb = F.conv2d(ones, T.view(Co, 1, 3, 3), None, padding=1)       # sum of T over the in-bounds taps at every p
    out[n] = y[0] + b.view(Co, 1, 1) - inb[0]
err = (out - ref).abs().max().item()
print("max |folded - reference| = %.3e" % err)
assert err < 1e-10
