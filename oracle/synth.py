"""Seeded synthetic checkpoints and inputs (oracle side; test infrastructure).

Both reference checkpoints are missing from the mount
(/root/reference/.MISSING_LARGE_BLOBS), so parity is proven on synthetic
weights stored in the two container formats the reference loads
(SURVEY.md App. E):

* ``gen_00000016.pt``  -> ``{'a': BDCN state_dict}``          (test.py:282-283)
* ``baseline_edge_16.pkl`` -> ``{'state_dict': ..., 'epoch': int}``  (test.py:294-295)

The key/shape tables below restate what the reference modules register
(bdcn_new.py:66-112, vgg16_c.py:11-39, models/RITnet_v2.py:15-29,32-88,91-121,
124-238, utils.py:983-1011,1039-1045); ``tests/golden/state_keys.json`` holds
the key->shape listing dumped from the real reference modules and the CPU test
suite checks these tables against it.

The reference's own initialisation is degenerate for parity work (edge map is a
constant 0.5 and logits reach 1e4), so weights are drawn
with analytic fan-in scaling; everything is a pure function of the seed.
"""
from collections import OrderedDict
import math

import numpy as np
import torch
import torch.nn.functional as F

H, W = 240, 320

VGG_CFG = [("conv1_1", 3, 64), ("conv1_2", 64, 64),
           ("conv2_1", 64, 128), ("conv2_2", 128, 128),
           ("conv3_1", 128, 256), ("conv3_2", 256, 256), ("conv3_3", 256, 256),
           ("conv4_1", 256, 512), ("conv4_2", 512, 512), ("conv4_3", 512, 512),
           ("conv5_1", 512, 512), ("conv5_2", 512, 512), ("conv5_3", 512, 512)]
BDCN_STAGE_BLOCKS = {1: 2, 2: 2, 3: 3, 4: 3, 5: 3}
BDCN_STAGE_CIN = {1: 64, 2: 128, 3: 256, 4: 512, 5: 512}
BDCN_UPSAMPLE = {"upsample_2": 4, "upsample_4": 8, "upsample_8": 16, "upsample_8_5": 16}


def bdcn_shapes():
    """key -> shape of BDCN().state_dict() (bdcn_new.py:66-112)."""
    s = OrderedDict()
    for name, ci, co in VGG_CFG:
        s[f"features.{name}.weight"] = (co, ci, 3, 3)
        s[f"features.{name}.bias"] = (co,)
    for st in range(1, 6):
        nb = BDCN_STAGE_BLOCKS[st]
        for j in range(1, nb + 1):
            cin = BDCN_STAGE_CIN[st]
            s[f"msblock{st}_{j}.conv.weight"] = (32, cin, 3, 3)
            s[f"msblock{st}_{j}.conv.bias"] = (32,)
            for c in ("conv1", "conv2", "conv3"):
                s[f"msblock{st}_{j}.{c}.weight"] = (32, 32, 3, 3)
                s[f"msblock{st}_{j}.{c}.bias"] = (32,)
        for j in range(1, nb + 1):
            s[f"conv{st}_{j}_down.weight"] = (21, 32, 1, 1)
            s[f"conv{st}_{j}_down.bias"] = (21,)
        for suf in ("", "_1"):
            s[f"score_dsn{st}{suf}.weight"] = (1, 21, 1, 1)
            s[f"score_dsn{st}{suf}.bias"] = (1,)
    for name, k in BDCN_UPSAMPLE.items():
        s[f"{name}.weight"] = (1, 1, k, k)
    s["fuse.weight"] = (1, 10, 1, 1)
    s["fuse.bias"] = (1,)
    return s


def esf_sizes(chz=32, growth=1.2, blks=4):
    """models/RITnet_v2.py:15-29 getSizes."""
    inter = [chz * (i + 1) for i in range(blks)]
    op = [int(growth * chz * (i + 1)) for i in range(blks)]
    ip = [chz] + [int(growth * chz * (i + 1)) for i in range(blks - 1)]
    skip = [a + b for a, b in zip(ip[::-1], inter[::-1])]
    return dict(inter=inter, op=op, ip=ip, skip=skip, dec_ip=op[::-1],
                dec_op=op[::-1][1:] + [chz])


def esf_shapes(setting):
    """key -> shape of DenseNet2D(setting).state_dict() (models/RITnet_v2.py:203-238)."""
    sz = esf_sizes()
    s = OrderedDict()

    def conv(name, co, ci, kh, kw, bias=True):
        s[name + ".weight"] = (co, ci, kh, kw)
        if bias:
            s[name + ".bias"] = (co,)

    def bn(name, c):
        s[name + ".weight"] = (c,)
        s[name + ".bias"] = (c,)
        s[name + ".running_mean"] = (c,)
        s[name + ".running_var"] = (c,)
        s[name + ".num_batches_tracked"] = ()

    in_c = 2 if setting["input_concat"] == 1 else 1
    conv("enc.head.conv1", 32, in_c, 3, 3)
    conv("enc.head.conv2", 32, 32, 3, 3)
    bn("enc.head.bn", 32)
    blocks = [("down_block1", sz["ip"][0], sz["inter"][0], sz["op"][0]),
              ("down_block2", sz["ip"][1], sz["inter"][1], sz["op"][1]),
              ("down_block3", sz["ip"][2], sz["inter"][2], sz["op"][2]),
              ("down_block4", sz["ip"][3], sz["inter"][3], sz["op"][3]),
              ("bottleneck", sz["op"][3], sz["inter"][3], sz["op"][3])]
    for name, ic, mc, oc in blocks:
        p = "enc." + name
        conv(p + ".conv1", mc, ic, 3, 3)
        conv(p + ".conv21", mc, ic + mc, 1, 1)
        conv(p + ".conv22", mc, mc, 3, 3)
        conv(p + ".conv31", mc, ic + 2 * mc, 1, 1)
        conv(p + ".conv32", mc, mc, 3, 3)
        conv(p + ".TD.conv", oc, ic + mc, 1, 1)
    dec_ip, dec_op = sz["dec_ip"], sz["dec_op"]
    if setting["add_edge"] == 1:                      # RITnet_v2.py:184-186
        dec_ip, dec_op = [306, 180, 100, 62], [180, 100, 62, 32]
    for i, name in enumerate(["up_block4", "up_block3", "up_block2", "up_block1"]):
        p = "dec." + name
        sk, ic, oc = sz["skip"][i], dec_ip[i], dec_op[i]
        conv(p + ".conv11", oc, sk + ic, 1, 1)
        conv(p + ".conv12", oc, oc, 3, 3)
        conv(p + ".conv21", oc, sk + ic + oc, 1, 1)
        conv(p + ".conv22", oc, oc, 3, 3)
    conv("dec.final.conv1", 32, 32, 3, 3)
    conv("dec.final.conv2", 3, 32, 3, 3)
    bn("dec.final.bn", 3)
    fc = setting["feature_channels"] * (2 if setting["add_edge"] == 1 else 1)
    if setting["add_seg"] == 1:                       # RITnet_v2.py:91-121,231-235
        sd = setting["style_dim"]
        conv("seg_encoder.model.0.conv", 64, 3, 7, 7)
        conv("seg_encoder.model.1.conv", 128, 64, 4, 4)
        conv("seg_encoder.model.2.conv", 256, 128, 4, 4)
        conv("seg_encoder.model.3.conv", 256, 256, 4, 4)
        conv("seg_encoder.model.4.conv", 256, 256, 4, 4)
        conv("seg_encoder.model.6", sd, 256, 1, 1)
        s["mlp.model.0.fc.weight"] = (256, sd); s["mlp.model.0.fc.bias"] = (256,)
        s["mlp.model.1.fc.weight"] = (256, 256); s["mlp.model.1.fc.bias"] = (256,)
        s["mlp.model.2.fc.weight"] = (2 * fc, 256); s["mlp.model.2.fc.bias"] = (2 * fc,)
    conv("elReg.c1", 128, fc, 2, 3)
    conv("elReg.c2", 128, 128, 3, 3)
    conv("elReg.c3", 32, 128, 3, 3, bias=False)
    s["elReg.l1.weight"] = (256, 480); s["elReg.l1.bias"] = (256,)
    s["elReg.l2.weight"] = (10, 256); s["elReg.l2.bias"] = (10,)
    return s


SETTINGS = {
    "baseline": dict(add_seg=0, seg_detach=0, add_edge=0, edge_thres=0, add_selayer=0,
                     generate_eyeball=0, feature_channels=153, style_dim=8,
                     input_concat=0, only_edge=0),
    "baseline_edge": dict(add_seg=0, seg_detach=0, add_edge=1, edge_thres=0, add_selayer=0,
                          generate_eyeball=0, feature_channels=153, style_dim=8,
                          input_concat=0, only_edge=0),
    "baseline_adain": dict(add_seg=1, seg_detach=0, add_edge=0, edge_thres=1, add_selayer=0,
                           generate_eyeball=0, feature_channels=153, style_dim=8,
                           input_concat=0, only_edge=0),
    "baseline_adain_edge": dict(add_seg=1, seg_detach=0, add_edge=1, edge_thres=1,
                                add_selayer=0, generate_eyeball=0, feature_channels=153,
                                style_dim=8, input_concat=0, only_edge=0),
    "baseline_input_concat": dict(add_seg=0, seg_detach=0, add_edge=0, edge_thres=0,
                                  add_selayer=0, generate_eyeball=0, feature_channels=153,
                                  style_dim=8, input_concat=1, only_edge=0),
    "baseline_only_edge": dict(add_seg=0, seg_detach=0, add_edge=0, edge_thres=0,
                               add_selayer=0, generate_eyeball=0, feature_channels=153,
                               style_dim=8, input_concat=0, only_edge=1),
}


def _bilinear_kernel(k):
    """bdcn_new.py:14-27 get_upsampling_weight restated for 1->1 channels."""
    factor = (k + 1) // 2
    center = factor - 1 if k % 2 == 1 else factor - 0.5
    og = np.arange(k, dtype=np.float64)
    f1 = 1 - np.abs(og - center) / factor
    return torch.from_numpy(np.outer(f1, f1)).float().reshape(1, 1, k, k)


def _fill(shapes, seed, gain_of):
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    for k, shp in shapes.items():
        if k.endswith("num_batches_tracked"):
            sd[k] = torch.tensor(100, dtype=torch.long)
        elif k.endswith("running_mean"):
            sd[k] = 0.2 * torch.randn(shp, generator=g)
        elif k.endswith("running_var"):
            sd[k] = 0.5 + torch.rand(shp, generator=g)
        elif len(shp) == 1 and (".bn." in k) and k.endswith("weight"):
            sd[k] = 0.75 + 0.5 * torch.rand(shp, generator=g)
        elif len(shp) == 1:                                   # biases
            sd[k] = 0.1 * torch.randn(shp, generator=g)
        else:
            fan_in = int(np.prod(shp[1:]))
            sd[k] = torch.randn(shp, generator=g) * (gain_of(k, shp) / math.sqrt(fan_in))
    return sd


def make_bdcn_state(seed=0):
    shapes = bdcn_shapes()

    def gain(k, shp):
        if "_down" in k:
            return 0.5
        if "score_dsn" in k:
            return 1.0
        return math.sqrt(2.0)

    sd = _fill(shapes, 1000 + seed, gain)
    g = torch.Generator().manual_seed(2000 + seed)
    for name, k in BDCN_UPSAMPLE.items():          # learnable, bilinear-initialised (F8)
        base = _bilinear_kernel(k)
        sd[name + ".weight"] = base * (1.0 + 0.1 * torch.randn(base.shape, generator=g))
    sd["fuse.weight"] = 0.08 + 0.03 * torch.randn((1, 10, 1, 1), generator=g)
    sd["fuse.bias"] = torch.tensor([-0.3])
    return sd


def make_esf_state(setting, seed=0):
    shapes = esf_shapes(setting)

    def gain(k, shp):
        if k.startswith("mlp.") or k.startswith("elReg.l"):
            return 1.0
        if len(shp) == 4 and shp[2] == 1 and shp[3] == 1:
            return 1.0                      # 1x1 pre-convs carry no activation
        return math.sqrt(2.0)

    sd = _fill(shapes, 3000 + seed, gain)
    return sd


def bdcn_checkpoint(seed=0):
    """Container of gen_00000016.pt (test.py:282-283)."""
    return {"a": make_bdcn_state(seed)}


def esf_checkpoint(setting, seed=0):
    """Container of baseline_edge_16.pkl (train.py:445-447, test.py:294-295)."""
    return {"state_dict": make_esf_state(setting, seed), "epoch": 16}


# ----------------------------------------------------------------------------- inputs

def synthetic_eye(idx, with_labels=True):
    """Structured synthetic eye: bright background, iris and pupil discs.

    Returns the 9-tuple the reference DataLoader yields per sample
    (CurriculumLib.py:139-166) as numpy arrays.
    """
    rng = np.random.RandomState(idx)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
    cx = W / 2 + rng.normal(0, 20)
    cy = H / 2 + rng.normal(0, 10)
    pcx = cx + rng.normal(0, 4)
    pcy = cy + rng.normal(0, 4)
    r_i = 60 + rng.normal(0, 4)
    r_p = 25 + rng.normal(0, 3)
    ang_i = rng.uniform(-0.5, 0.5)
    ai, bi = r_i, r_i * rng.uniform(0.8, 1.0)
    ap, bp = r_p, r_p * rng.uniform(0.8, 1.0)

    def inside(cx_, cy_, a, b, th):
        X = (xx - cx_) * np.cos(th) + (yy - cy_) * np.sin(th)
        Y = -(xx - cx_) * np.sin(th) + (yy - cy_) * np.cos(th)
        return (X / a) ** 2 + (Y / b) ** 2 <= 1

    img = np.full((H, W), 200.0)
    label = np.zeros((H, W), np.int64)
    m_i = inside(cx, cy, ai, bi, ang_i)
    m_p = inside(pcx, pcy, ap, bp, ang_i)
    img[m_i] = 140.0
    label[m_i] = 1
    img[m_p] = 50.0
    label[m_p] = 2
    img += rng.normal(0, 8, img.shape)
    img = np.clip(img, 0, 255)
    img_u8 = img.astype(np.uint8)
    z = (img_u8 - img_u8.mean()) / img_u8.std()
    elnorm = np.array([[2 * cx / W - 1, 2 * cy / H - 1, 2 * ai / W, 2 * bi / H, ang_i],
                       [2 * pcx / W - 1, 2 * pcy / H - 1, 2 * ap / W, 2 * bp / H, ang_i]])
    return dict(img=z.astype(np.float32)[None], img_u8=img_u8, label=label,
                spatW=np.ones((H, W), np.float32), distMap=np.zeros((3, H, W), np.float32),
                pupil_center=np.array([pcx, pcy], np.float32),
                iris_center=np.array([cx, cy], np.float32),
                elNorm=elnorm.astype(np.float32), cond=np.zeros(4, np.float32),
                imInfo=np.array([idx, 0, 0], np.int64))


def synthetic_eye_batch(start, count):
    items = [synthetic_eye(start + i) for i in range(count)]
    out = {}
    for k in items[0]:
        out[k] = np.stack([it[k] for it in items], 0)
    return out


def randn_frames(b, seed=0):
    """seeded z-score-like noise frames [B,1,240,320]."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(b, 1, H, W, generator=g)


def evaluate_style_labels(b):
    """Fake labels holding all three classes (evaluate.py:118-120)."""
    lab = torch.zeros((b, H, W), dtype=torch.long)
    lab[:, 0, 2] = 1
    lab[:, 2, 2] = 2
    return lab


def _blur(x, k=9):
    """Box blur of a [B,C,H,W] float tensor (deterministic helper for the synthetic loss inputs)."""
    return F.avg_pool2d(x, k, 1, k // 2, count_include_pad=True)


def smooth_logits(label, seed=0):
    """Plausible segmentation logits [B,3,H,W] for labels [B,H,W]: 4 * one-hot + smooth noise, so the
    softmax is neither saturated nor uniform (inputs of the loss-slot fixtures, tests/golden/loss.npz)."""
    lab = torch.as_tensor(np.asarray(label)).long()
    g = torch.Generator().manual_seed(seed)
    onehot = F.one_hot(lab, 3).permute(0, 3, 1, 2).float()
    return 4.0 * onehot + 2.0 * _blur(torch.randn(onehot.shape, generator=g)) + 0.3 * torch.randn(onehot.shape, generator=g)


def loss_maps(label, seed=0):
    """Synthetic spatial weights [B,H,W] (>= 1) and signed distance-like maps [B,3,H,W] for the loss slot."""
    lab = torch.as_tensor(np.asarray(label))
    g = torch.Generator().manual_seed(seed)
    B = lab.shape[0]
    sw = 1.0 + 6.0 * _blur(torch.randn((B, 1, H, W), generator=g)).abs()[:, 0]
    dm = 20.0 * _blur(torch.randn((B, 3, H, W), generator=g), 15)
    return sw, dm
