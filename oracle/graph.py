"""Torch-fp32 functional restatement of the reference hot path (CPU oracle).

TEST INFRASTRUCTURE ONLY - see oracle/__init__.py.  Every function cites the
reference lines it follows (paths relative to /root/reference).  Nothing here
shares code with the product: the product is the CUDA library behind
include/egn.h.

All functions take plain ``state_dict`` mappings (name -> fp32 tensor) in the
reference's key layout and run on CPU in fp32.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

LRELU = 0.01  # F.leaky_relu default slope, used as actfunc (RITnet_v2.py:207)


# ----------------------------------------------------------------------------- BDCN

def vgg16c_features(sd, x, taps=None):
    """vgg16_c.py:65-88 - 13 conv3x3+ReLU, ceil-mode max pools, dilated stage 5."""
    def c(name, t, dil=1):
        return F.relu(F.conv2d(t, sd[f"features.{name}.weight"], sd[f"features.{name}.bias"],
                               padding=dil, dilation=dil))
    c11 = c("conv1_1", x); c12 = c("conv1_2", c11)
    p1 = F.max_pool2d(c12, 2, 2, ceil_mode=True)
    c21 = c("conv2_1", p1); c22 = c("conv2_2", c21)
    p2 = F.max_pool2d(c22, 2, 2, ceil_mode=True)
    c31 = c("conv3_1", p2); c32 = c("conv3_2", c31); c33 = c("conv3_3", c32)
    p3 = F.max_pool2d(c33, 2, 2, ceil_mode=True)
    c41 = c("conv4_1", p3); c42 = c("conv4_2", c41); c43 = c("conv4_3", c42)
    p4 = F.max_pool2d(c43, 2, 1, ceil_mode=True)          # vgg16_c.py:34 stride 1
    c51 = c("conv5_1", p4, 2); c52 = c("conv5_2", c51, 2); c53 = c("conv5_3", c52, 2)
    return [c11, c12, c21, c22, c31, c32, c33, c41, c42, c43, c51, c52, c53]


def msblock(sd, prefix, x, rate=4):
    """bdcn_new.py:49-55."""
    o = F.relu(F.conv2d(x, sd[prefix + ".conv.weight"], sd[prefix + ".conv.bias"], padding=1))
    out = o
    for i, name in enumerate(("conv1", "conv2", "conv3")):
        d = rate * (i + 1)
        out = out + F.relu(F.conv2d(o, sd[f"{prefix}.{name}.weight"], sd[f"{prefix}.{name}.bias"],
                                    padding=d, dilation=d))
    return out


def bdcn_forward(sd, x, return_all=False, taps=None):
    """bdcn_new.py:116-191.  x: [B,3,H,W].  Returns the fuse map (index [-1])."""
    B, _, H, W = x.shape
    feats = vgg16c_features(sd, x)
    nblk = {1: 2, 2: 2, 3: 3, 4: 3, 5: 3}
    ups = {2: ("upsample_2", 2, 1), 3: ("upsample_4", 4, 2),
           4: ("upsample_8", 8, 4), 5: ("upsample_8_5", 8, 0)}
    s_a, s_b = {}, {}
    fi = 0
    for st in range(1, 6):
        acc = None
        for j in range(1, nblk[st] + 1):
            ms = msblock(sd, f"msblock{st}_{j}", feats[fi]); fi += 1
            if taps is not None:
                taps[f"msblock{st}_{j}"] = ms
            d = F.conv2d(ms, sd[f"conv{st}_{j}_down.weight"], sd[f"conv{st}_{j}_down.bias"])
            acc = d if acc is None else acc + d
        a = F.conv2d(acc, sd[f"score_dsn{st}.weight"], sd[f"score_dsn{st}.bias"])
        b = F.conv2d(acc, sd[f"score_dsn{st}_1.weight"], sd[f"score_dsn{st}_1.bias"])
        if st > 1:
            name, stride, off = ups[st]
            w = sd[name + ".weight"]
            a = F.conv_transpose2d(a, w, stride=stride)[:, :, off:off + H, off:off + W]
            b = F.conv_transpose2d(b, w, stride=stride)[:, :, off:off + H, off:off + W]
        s_a[st], s_b[st] = a, b
    p1 = [s_a[k] + sum(s_a[j] for j in range(1, k)) if k > 1 else s_a[1] for k in range(1, 6)]
    p2 = [s_b[k] + sum(s_b[j] for j in range(k + 1, 6)) if k < 5 else s_b[5] for k in range(1, 6)]
    allp = p1 + p2
    fuse = F.conv2d(torch.cat(allp, 1), sd["fuse.weight"], sd["fuse.bias"])
    if taps is not None:
        taps["features"] = feats
    if return_all:
        return [torch.sigmoid(p) for p in allp] + [torch.sigmoid(fuse)]
    return torch.sigmoid(fuse)


def calc_edge(sd, img, edge_thres=0):
    """utils.py:645-656: replicate grey to 3 channels, take BDCN()[-1], optional threshold."""
    e = bdcn_forward(sd, torch.cat((img, img, img), 1))
    if edge_thres == 1:
        e = torch.where(e >= 0.1, torch.ones_like(e), e)
    return e


# ----------------------------------------------------------------------------- ESF-Net

def _conv(sd, name, x, pad=0):
    return F.conv2d(x, sd[name + ".weight"], sd.get(name + ".bias"), padding=pad)


def _bn_eval(sd, name, x):
    return F.batch_norm(x, sd[name + ".running_mean"], sd[name + ".running_var"],
                        sd[name + ".weight"], sd[name + ".bias"], False, 0.0, 1e-5)


def conv_block(sd, p, x):
    """utils.py:1046-1050 convBlock: conv-act-conv-act-BN (BN after the activation)."""
    x = F.leaky_relu(_conv(sd, p + ".conv1", x, 1), LRELU)
    x = F.leaky_relu(_conv(sd, p + ".conv2", x, 1), LRELU)
    return _bn_eval(sd, p + ".bn", x)


def down_block(sd, p, x, pool, taps=None):
    """RITnet_v2.py:59-66 + Transition_down :40-44; norm = InstanceNorm2d (F4)."""
    x1 = F.leaky_relu(_conv(sd, p + ".conv1", F.instance_norm(x, eps=1e-5), 1), LRELU)
    x21 = torch.cat([x, x1], 1)
    x22 = F.leaky_relu(_conv(sd, p + ".conv22", _conv(sd, p + ".conv21", x21), 1), LRELU)
    x31 = torch.cat([x21, x22], 1)
    out = F.leaky_relu(_conv(sd, p + ".conv32", _conv(sd, p + ".conv31", x31), 1), LRELU)
    out = torch.cat([out, x], 1)
    t = F.leaky_relu(F.instance_norm(out, eps=1e-5), LRELU)
    t = _conv(sd, p + ".TD.conv", t)
    if pool:
        t = F.avg_pool2d(t, 2)
    if taps is not None:
        taps[p + ".x1"] = x1; taps[p + ".x22"] = x22; taps[p + ".skip"] = out; taps[p + ".td"] = t
    return out, t


def encoder(sd, x, taps=None, tag=""):
    """RITnet_v2.py:167-174."""
    x = conv_block(sd, "enc.head", x)
    if taps is not None:
        taps["enc.head" + tag] = x
    t = None if taps is None else {}
    s1, x = down_block(sd, "enc.down_block1", x, True, t)
    s2, x = down_block(sd, "enc.down_block2", x, True, t)
    s3, x = down_block(sd, "enc.down_block3", x, True, t)
    s4, x = down_block(sd, "enc.down_block4", x, True, t)
    _, x = down_block(sd, "enc.bottleneck", x, False, t)
    if taps is not None:
        for k, v in t.items():
            taps[k + tag] = v
    return s4, s3, s2, s1, x


def up_block(sd, p, skip, x):
    """RITnet_v2.py:79-88."""
    x = F.interpolate(x, mode="bilinear", align_corners=False, scale_factor=2)
    x = torch.cat([x, skip], 1)
    x1 = F.leaky_relu(_conv(sd, p + ".conv12", _conv(sd, p + ".conv11", x), 1), LRELU)
    x21 = torch.cat([x, x1], 1)
    return F.leaky_relu(_conv(sd, p + ".conv22", _conv(sd, p + ".conv21", x21), 1), LRELU)


def decoder(sd, s4, s3, s2, s1, x, taps=None):
    """RITnet_v2.py:194-200."""
    x = up_block(sd, "dec.up_block4", s4, x)
    if taps is not None: taps["dec.up4"] = x
    x = up_block(sd, "dec.up_block3", s3, x)
    if taps is not None: taps["dec.up3"] = x
    x = up_block(sd, "dec.up_block2", s2, x)
    if taps is not None: taps["dec.up2"] = x
    x = up_block(sd, "dec.up_block1", s1, x)
    if taps is not None: taps["dec.up1"] = x
    return conv_block(sd, "dec.final", x)


def regression_module(sd, x):
    """utils.py:1013-1037."""
    B = x.shape[0]
    x = F.leaky_relu(_conv(sd, "elReg.c1", x), LRELU)
    x = F.avg_pool2d(x, 2)
    x = F.leaky_relu(_conv(sd, "elReg.c2", x), LRELU)
    x = F.leaky_relu(F.conv2d(x, sd["elReg.c3.weight"]), LRELU)
    x = x.reshape(B, -1)
    x = F.linear(torch.selu(F.linear(x, sd["elReg.l1.weight"], sd["elReg.l1.bias"])),
                 sd["elReg.l2.weight"], sd["elReg.l2.bias"])
    return torch.cat([torch.tanh(x[:, 0:2]), torch.sigmoid(x[:, 2:4]), x[:, 4:5],
                      torch.tanh(x[:, 5:7]), torch.sigmoid(x[:, 7:9]), x[:, 9:10]], 1)


def style_encoder(sd, x):
    """RITnet_v2.py:91-106 with Conv2dBlock utils.py:1093-1149 (reflect pad, ReLU, no norm)."""
    def blk(i, t, k, s, p):
        t = F.pad(t, (p, p, p, p), mode="reflect")
        return F.relu(F.conv2d(t, sd[f"seg_encoder.model.{i}.conv.weight"],
                               sd[f"seg_encoder.model.{i}.conv.bias"], stride=s))
    x = blk(0, x, 7, 1, 3)
    for i in (1, 2, 3, 4):
        x = blk(i, x, 4, 2, 1)
    x = F.adaptive_avg_pool2d(x, 1)
    return F.conv2d(x, sd["seg_encoder.model.6.weight"], sd["seg_encoder.model.6.bias"])


def mlp(sd, x):
    """RITnet_v2.py:109-121 with LinearBlock utils.py:1051-1091."""
    x = x.view(x.size(0), -1)
    x = F.relu(F.linear(x, sd["mlp.model.0.fc.weight"], sd["mlp.model.0.fc.bias"]))
    x = F.relu(F.linear(x, sd["mlp.model.1.fc.weight"], sd["mlp.model.1.fc.bias"]))
    return F.linear(x, sd["mlp.model.2.fc.weight"], sd["mlp.model.2.fc.bias"])


def seg2pt(op, temperature=4.0):
    """loss.py:16-46 get_seg2ptLoss + utils.py:27-60 create_meshgrid (prediction half only).

    op [B,H,W] -> [B,2] expectation of linspace(-1,1) grids under softmax(T*op)."""
    B, H, W = op.shape
    wt = F.softmax(op.reshape(B, -1) * temperature, dim=1).view(B, H, W)
    xs = torch.linspace(-1, 1, W)
    ys = torch.linspace(-1, 1, H)
    xpos = (wt.sum(1) * xs).sum(-1)
    ypos = (wt.sum(2) * ys).sum(-1)
    return torch.stack([xpos, ypos], 1)


def seg2pt_exact(op, temperature=4.0):
    """Same as seg2pt but with the reference's exact summation layout (flattened grids)."""
    B, H, W = op.shape
    wt = F.softmax(op.reshape(B, -1) * temperature, dim=1)
    xs = torch.linspace(-1, 1, W)
    ys = torch.linspace(-1, 1, H)
    xloc = xs.view(1, W).expand(H, W).reshape(-1)
    yloc = ys.view(H, 1).expand(H, W).reshape(-1)
    return torch.stack([(wt * xloc).sum(-1), (wt * yloc).sum(-1)], 1)


def esf_forward(sd, setting, x, x_edge, has_mask=None, taps=None):
    """RITnet_v2.py:261-354 (inference-relevant outputs; the training loss slot is not restated).

    has_mask: [B] bool, True where a GT mask exists, i.e. ``1 - cond[:,1]``
    (RITnet_v2.py:383,393-408); ``None`` = all True.
    Returns dict(op, elPred, latent, elOut)."""
    B = x.shape[0]
    if setting["only_edge"] == 1:
        x = x_edge
    if setting["input_concat"] == 1:
        x = torch.cat((x, x_edge), 1)
    s4, s3, s2, s1, xb = encoder(sd, x, taps)
    latent = xb.flatten(2).mean(-1)
    if setting["add_edge"] == 1:
        _, _, _, _, xe = encoder(sd, x_edge, taps, tag="@edge")
        xb = torch.cat((xb, xe), 1)
    op = decoder(sd, s4, s3, s2, s1, xb, taps)
    xr = xb
    if setting["add_seg"] == 1:
        enc = style_encoder(sd, F.softmax(op, 1))
        ad = mlp(sd, enc).view(B, 2, -1)
        C = xb.shape[1]
        flat = xb.view(B, C, -1)
        std = (flat.var(dim=2) + 1e-5).sqrt().view(B, C, 1, 1)     # unbiased (RITnet_v2.py:256)
        mean = flat.mean(dim=2).view(B, C, 1, 1)
        xr = (xb - mean) / std * ad[:, 0].view(B, C, 1, 1) + ad[:, 1].view(B, C, 1, 1)
        if taps is not None:
            taps["adain"] = ad
    el = regression_module(sd, xr)
    pup = seg2pt_exact(op[:, 2])
    if has_mask is None or bool(torch.as_tensor(has_mask).any()):
        iri = seg2pt_exact(-op[:, 0])
    else:
        iri = el[:, 5:7].clone()                                    # RITnet_v2.py:403-408
    elpred = torch.cat([iri, el[:, 2:5], pup, el[:, 7:10]], 1)       # RITnet_v2.py:334-335
    if taps is not None:
        taps["bottleneck_cat"] = xb
    return dict(op=op, elPred=elpred, latent=latent, elOut=el)


def get_predictions(op):
    """utils.py:65-81: argmax over channels of raw logits, first index on ties."""
    return op.max(1)[1]


# ----------------------------------------------------------------------------- loss slot

def norm_pts(pts, sz):
    """utils.py:627-634 normPts: pixel (x, y) -> [-1, 1] with x / W, y / H."""
    p = torch.as_tensor(pts, dtype=torch.float32).clone().reshape(-1, 2)
    p[:, 0] = 2 * (p[:, 0] / sz[1]) - 1
    p[:, 1] = 2 * (p[:, 1] / sz[0]) - 1
    return p


def surface_loss(op_i, dist_i):
    """loss.py:91-97 SurfaceLoss for one sample: mean over channels of the pixel mean of
    softmax(op) * distmap."""
    p = torch.softmax(op_i, 0).flatten(1)
    return (p * dist_i.flatten(1)).mean(1).mean()


def wce_loss(op_i, target_i, spat_i):
    """loss.py:125-137 wCE: mean(spatWts * cross_entropy(...)) where cross_entropy is the scalar
    mean over pixels (ignore_index is the single class absent from the target, which no pixel
    carries, so nothing is ignored)."""
    ce = F.cross_entropy(op_i.flatten(1).unsqueeze(0), target_i.flatten().unsqueeze(0).long())
    return (spat_i.flatten() * ce).mean()


def gdice_loss(op_i, target_i):
    """loss.py:99-123 GDiceLoss for one sample: class weights 1 / clamp(count^2, 1e-5), zero for
    classes absent from the target; 1 - clamp(2 * sum(w * p.t) / sum(w * (p + t)), 1e-5)."""
    C = op_i.shape[0]
    p = torch.softmax(op_i, 0).flatten(1)
    t = torch.stack([(target_i.flatten() == c).to(p.dtype) for c in range(C)])
    cnt = t.sum(1)
    w = 1.0 / (cnt ** 2).clamp(1e-5)
    w = torch.where(cnt > 0, w, torch.zeros_like(w))
    dice = 2.0 * (w * (p * t).sum(1)).sum() / (w * (p + t).sum(1)).sum()
    return 1 - dice.clamp(1e-5)


def all_loss(op, el_out, target, pupil_center, el_norm, spat_w, dist_map, cond, alpha):
    """models/RITnet_v2.py:372-440 get_allLoss (with loss.py:48-89 get_segLoss / get_ptLoss):
    returns (total_loss scalar, pred_c_seg [B,2,2] iris first)."""
    B, C, H, W = op.shape
    mask = (1 - cond[:, 1]).to(torch.float32)
    gt_pup = norm_pts(pupil_center, (H, W))
    pup = seg2pt_exact(op[:, 2])
    l_pup = (pup - gt_pup).abs()
    if mask.sum() > 0:
        iri = seg2pt_exact(-op[:, 0])
        l_iri = (iri - el_norm[:, 0, :2]).abs()
        tmp = torch.stack([mask, mask], 1)
        l_iri = (l_iri * tmp).sum() / tmp.sum()
    else:
        l_iri = 0.0
        iri = el_out[:, 5:7].clone()
    l_seg2pt = 0.5 * l_pup.mean() + 0.5 * l_iri
    seg = [alpha * surface_loss(op[i], dist_map[i]) + (1 - alpha) * gdice_loss(op[i], target[i]) +
           wce_loss(op[i], target[i], spat_w[i]) for i in range(B) if mask[i] == 1]
    l_seg = torch.stack(seg).sum() / mask.sum() if seg else 0.0

    def pt_loss(a, b, c):
        v = [(a[i] - b[i]).abs().mean() for i in range(B) if c[i] == 1]
        return torch.stack(v).sum() / c.sum() if v else 0.0

    l_pt = pt_loss(el_out[:, 5:7], gt_pup, 1 - mask)
    l_el = pt_loss(el_out, el_norm.reshape(-1, 10), mask)
    total = l_seg2pt + 20 * l_seg + 10 * (l_pt + l_el)
    return torch.as_tensor(total, dtype=torch.float32), torch.stack([iri, pup], 1)


# ----------------------------------------------------------------------------- metrics

def seg_metrics(y_true, y_pred, cond):
    """utils.py:120-150 getSeg_metrics without sklearn: per-sample Jaccard for the classes
    present in GT, NaN elsewhere, nanmean over valid samples."""
    y_true = np.asarray(y_true); y_pred = np.asarray(y_pred)
    cond = np.asarray(cond).astype(bool)
    B = y_true.shape[0]
    scores = np.full((B, 3), np.nan)
    for i in range(B):
        if cond[i]:
            continue
        for c in np.unique(y_true[i]):
            t = y_true[i] == c
            p = y_pred[i] == c
            scores[i, int(c)] = (t & p).sum() / float((t | p).sum())
    clean = scores[~cond]
    if len(clean) == 0:
        return np.nan, np.full(3, np.nan), scores
    with np.errstate(all="ignore"):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            per = np.nanmean(clean, 0)
            return np.nanmean(per), per, scores


def unnorm_pts(pts, sz):
    """utils.py:636-643."""
    p = np.array(pts, dtype=np.float64).reshape(-1, 2).copy()
    p[:, 0] = 0.5 * sz[1] * (p[:, 0] + 1)
    p[:, 1] = 0.5 * sz[0] * (p[:, 1] + 1)
    return p.reshape(np.shape(pts))


def point_metric(y_true, y_pred, cond, sz, do_unnorm=True):
    """utils.py:152-162 getPoint_metric."""
    if do_unnorm:
        y_pred = unnorm_pts(y_pred, sz)
    flag = (~np.asarray(cond).astype(bool)).astype(np.float64)
    d = flag * np.sqrt(((np.asarray(y_true, np.float64) - y_pred) ** 2).sum(-1))
    return (d.sum() / flag.sum() if flag.any() else np.nan), d


# ----------------------------------------------------------------------------- ellipse geometry

def _rot(t):
    c, s = math.cos(t), math.sin(t)
    return np.array([[c, -s, 0.0], [s, c, 0.0], [0, 0, 1.0]])


def _trans(tx, ty):
    return np.array([[1.0, 0, tx], [0, 1.0, ty], [0, 0, 1.0]])


def ellipse_param2mat(p):
    """helperfunctions.py:25-33 my_ellipse.param2mat."""
    cx, cy, a, b, th = [float(v) for v in p[:5]]
    Hr, Ht = _rot(-th), _trans(-cx, -cy)
    Q = np.diag([1 / a ** 2, 1 / b ** 2, -1.0])
    return Ht.T @ Hr.T @ Q @ Hr @ Ht


def ellipse_mat2param(m):
    """helperfunctions.py:50-63,102-123 (mat2param, recover_theta, recover_C)."""
    a, b, c, d, e = m[0, 0], 2 * m[0, 1], m[1, 1], 2 * m[0, 2], 2 * m[1, 2]
    EPS = 1e-40
    if abs(b) <= EPS and a <= c:
        th = 0.0
    elif abs(b) <= EPS and a > c:
        th = math.pi / 2
    else:
        th = 0.5 * math.atan2(b, (a - c))
    den = b ** 2 - 4 * a * c
    tx = (2 * c * d - b * e) / den
    ty = (2 * a * e - b * d) / den
    Hr, Ht = _rot(th), _trans(tx, ty)
    mn = Hr.T @ Ht.T @ m @ Ht @ Hr
    major = math.sqrt(1 / mn[0, 0])
    minor = math.sqrt(1 / mn[1, 1])
    return np.array([tx, ty, major, minor, th, math.pi * major * minor])


def ellipse_transform(p, Hm):
    """helperfunctions.py:124-129 my_ellipse.transform -> parameters only."""
    Hi = np.linalg.inv(Hm)
    return ellipse_mat2param(Hi.T @ ellipse_param2mat(p) @ Hi)


def ellipse_norm_to_px(p, H=240, W=320):
    """evaluate.py:141-146: normalised -> pixel ellipse, first five parameters."""
    Hm = np.array([[W / 2, 0, W / 2], [0, H / 2, H / 2], [0, 0, 1.0]])
    return ellipse_transform(p, Hm)[:-1]


def ell_iou(seg, el_px_deg):
    """utils.py:176-204 calc_ell_iou(nor=False, angle_nor=True): IoU between a boolean mask and
    the raster of a pixel-space ellipse (angle in degrees, pi := 3.14159) drawn on the [-1,1] grid."""
    Hh, Ww = seg.shape
    el = np.array(el_px_deg, dtype=np.float64).copy()
    el[4] = el[4] / 180.0 * 3.14159
    Hm = np.array([[2 / Ww, 0, -1], [0, 2 / Hh, -1], [0, 0, 1.0]])
    e = ellipse_transform(el, Hm)[:-1]
    # the reference mixes a float32 torch meshgrid with float64 numpy scalars, so the raster
    # arithmetic runs in float32 (utils.py:191-196); keep that to stay bit-exact on the mask
    xs = torch.linspace(-1, 1, Ww).view(1, Ww).expand(Hh, Ww)
    ys = torch.linspace(-1, 1, Hh).view(Hh, 1).expand(Hh, Ww)
    c, s = float(np.cos(e[4])), float(np.sin(e[4]))
    X = (xs - float(e[0])) * c + (ys - float(e[1])) * s
    Y = -(xs - float(e[0])) * s + (ys - float(e[1])) * c
    m = (((X / float(e[2])) ** 2 + (Y / float(e[3])) ** 2 - 1) <= 0).numpy()
    seg = np.asarray(seg).astype(bool)
    inter = np.float32((seg & m).sum())
    union = np.float32(np.float32(seg.sum()) + np.float32(m.sum())) - inter
    with np.errstate(all="ignore"):
        return float(np.float32(inter) / np.float32(union))     # float32 like the torch sums


def refine_ellipse(seg, ell_px):
    """utils.py:450-486 search_proper_parameter_iou_for_our_data: coordinate descent on
    (a, b, theta_deg) maximising mask IoU; steps start at 1 and shrink x0.8; <=40 sweeps."""
    center = [float(ell_px[0]), float(ell_px[1])]
    ans = [float(ell_px[2]), float(ell_px[3]), float(ell_px[4]) * 180.0 / 3.14159]
    rt = ell_iou(seg, center + ans)
    now = list(ans)
    d = [1.0, 1.0, 1.0]
    for _ in range(40):
        flag = False
        for j in range(3):
            now[j] -= d[j]
            if ell_iou(seg, center + now) > rt:
                flag = True
                continue
            now[j] += 2.0 * d[j]
            if ell_iou(seg, center + now) > rt:
                flag = True
                continue
            now[j] -= d[j]
            d[j] *= 0.8
        sc = ell_iou(seg, center + now)
        if sc > rt:
            rt = sc
        if not flag:
            break
    out = np.array(center + now)
    out[4] = out[4] / 180.0 * 3.14159
    return out


def preprocess_frame_u8(img_u8):
    """evaluate.py:102-103 / CurriculumLib.py:139-140: per-frame z-score (population std)."""
    img = np.asarray(img_u8).astype(np.float64)
    return ((img - img.mean()) / img.std()).astype(np.float32)
