"""Parameter name -> shape tables of the two reference modules, used to declare look-alike
nn.Modules whose state_dict key sets match the reference exactly (strict loading, test.py:283,295).

BDCN: bdcn_new.py:66-112 + vgg16_c.py:11-39.  DenseNet2D: models/RITnet_v2.py:15-29,203-238,
utils.py:983-1011,1039-1045.  tests/test_boundary_cpu.py checks these against the key listing dumped
from the real reference (tests/golden/state_keys.json)."""
from collections import OrderedDict

_VGG = [("conv1_1", 3, 64), ("conv1_2", 64, 64), ("conv2_1", 64, 128), ("conv2_2", 128, 128),
        ("conv3_1", 128, 256), ("conv3_2", 256, 256), ("conv3_3", 256, 256), ("conv4_1", 256, 512),
        ("conv4_2", 512, 512), ("conv4_3", 512, 512), ("conv5_1", 512, 512), ("conv5_2", 512, 512),
        ("conv5_3", 512, 512)]
_STAGES = [(1, 2, 64), (2, 2, 128), (3, 3, 256), (4, 3, 512), (5, 3, 512)]   # stage, blocks, cin


def _wb(table, name, shape, bias=True):
    table[name + ".weight"] = tuple(shape)
    if bias:
        table[name + ".bias"] = (shape[0],)


def _bn(table, name, c):
    for k in ("weight", "bias", "running_mean", "running_var"):
        table["%s.%s" % (name, k)] = (c,)
    table[name + ".num_batches_tracked"] = ()


def bdcn_param_shapes():
    t = OrderedDict()
    for name, ci, co in _VGG:
        _wb(t, "features." + name, (co, ci, 3, 3))
    for st, nb, cin in _STAGES:
        for j in range(1, nb + 1):
            _wb(t, "msblock%d_%d.conv" % (st, j), (32, cin, 3, 3))
            for c in ("conv1", "conv2", "conv3"):
                _wb(t, "msblock%d_%d.%s" % (st, j, c), (32, 32, 3, 3))
        for j in range(1, nb + 1):
            _wb(t, "conv%d_%d_down" % (st, j), (21, 32, 1, 1))
        _wb(t, "score_dsn%d" % st, (1, 21, 1, 1))
        _wb(t, "score_dsn%d_1" % st, (1, 21, 1, 1))
    for name, k in (("upsample_2", 4), ("upsample_4", 8), ("upsample_8", 16), ("upsample_8_5", 16)):
        t[name + ".weight"] = (1, 1, k, k)
    _wb(t, "fuse", (1, 10, 1, 1))
    return t


def esf_sizes(chz=32, growth=1.2, blks=4):
    inter = [chz * (i + 1) for i in range(blks)]
    op = [int(growth * chz * (i + 1)) for i in range(blks)]
    ip = [chz] + op[:-1]
    return {"enc": {"inter": inter, "ip": ip, "op": op},
            "dec": {"skip": [a + b for a, b in zip(ip[::-1], inter[::-1])], "ip": op[::-1],
                    "op": op[::-1][1:] + [chz]}}


def esf_param_shapes(setting):
    sz = esf_sizes()
    t = OrderedDict()
    _wb(t, "enc.head.conv1", (32, 2 if setting["input_concat"] == 1 else 1, 3, 3))
    _wb(t, "enc.head.conv2", (32, 32, 3, 3))
    _bn(t, "enc.head.bn", 32)
    enc = sz["enc"]
    blocks = [("down_block%d" % (i + 1), enc["ip"][i], enc["inter"][i], enc["op"][i]) for i in range(4)]
    blocks.append(("bottleneck", enc["op"][3], enc["inter"][3], enc["op"][3]))
    for name, ic, mc, oc in blocks:
        p = "enc.%s." % name
        _wb(t, p + "conv1", (mc, ic, 3, 3))
        _wb(t, p + "conv21", (mc, ic + mc, 1, 1))
        _wb(t, p + "conv22", (mc, mc, 3, 3))
        _wb(t, p + "conv31", (mc, ic + 2 * mc, 1, 1))
        _wb(t, p + "conv32", (mc, mc, 3, 3))
        _wb(t, p + "TD.conv", (oc, ic + mc, 1, 1))
    dip, dop = sz["dec"]["ip"], sz["dec"]["op"]
    if setting["add_edge"] == 1:
        dip, dop = [306, 180, 100, 62], [180, 100, 62, 32]
    for i, name in enumerate(("up_block4", "up_block3", "up_block2", "up_block1")):
        p = "dec.%s." % name
        sk, ic, oc = sz["dec"]["skip"][i], dip[i], dop[i]
        _wb(t, p + "conv11", (oc, sk + ic, 1, 1))
        _wb(t, p + "conv12", (oc, oc, 3, 3))
        _wb(t, p + "conv21", (oc, sk + ic + oc, 1, 1))
        _wb(t, p + "conv22", (oc, oc, 3, 3))
    _wb(t, "dec.final.conv1", (32, 32, 3, 3))
    _wb(t, "dec.final.conv2", (3, 32, 3, 3))
    _bn(t, "dec.final.bn", 3)
    fc = setting["feature_channels"] * (2 if setting["add_edge"] == 1 else 1)
    if setting["add_seg"] == 1:
        sd = setting["style_dim"]
        _wb(t, "seg_encoder.model.0.conv", (64, 3, 7, 7))
        _wb(t, "seg_encoder.model.1.conv", (128, 64, 4, 4))
        _wb(t, "seg_encoder.model.2.conv", (256, 128, 4, 4))
        _wb(t, "seg_encoder.model.3.conv", (256, 256, 4, 4))
        _wb(t, "seg_encoder.model.4.conv", (256, 256, 4, 4))
        _wb(t, "seg_encoder.model.6", (sd, 256, 1, 1))
        _wb(t, "mlp.model.0.fc", (256, sd))
        _wb(t, "mlp.model.1.fc", (256, 256))
        _wb(t, "mlp.model.2.fc", (2 * fc, 256))
    _wb(t, "elReg.c1", (128, fc, 2, 3))
    _wb(t, "elReg.c2", (128, 128, 3, 3))
    _wb(t, "elReg.c3", (32, 128, 3, 3), bias=False)
    _wb(t, "elReg.l1", (256, 480))
    _wb(t, "elReg.l2", (10, 256))
    return t
