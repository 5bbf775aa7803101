"""Imports the UNMODIFIED reference from /root/reference (build container only).

TEST INFRASTRUCTURE ONLY.  /root/reference does not exist on the GPU box, so
nothing in ``-m gpu`` tests, ``smoke()`` or ``bench.py`` may import this module;
it is used by ``oracle/make_golden.py`` to produce ``tests/golden/*`` and by the
optional ``tests/test_oracle_vs_reference.py`` (skipped when the mount is absent).

Shims: numpy-2 aliases, stub ``skimage``/``matplotlib``,
``resource.setrlimit`` no-op, ``.cuda()`` neutralised on CPU-only hosts.
"""
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("EGN_REFERENCE_DIR", "/root/reference")


def available():
    return os.path.isdir(REF) and os.path.isfile(os.path.join(REF, "bdcn_new.py"))


_installed = False


def install_shims():
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError("reference mount %s not present" % REF)
    for name, typ in (("int", int), ("bool", bool), ("float", float)):
        if not hasattr(np, name):
            setattr(np, name, typ)
    empty = (np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64))
    sk = types.ModuleType("skimage"); skd = types.ModuleType("skimage.draw")
    skd.ellipse_perimeter = lambda *a, **k: empty
    skd.disk = lambda *a, **k: empty
    skd.ellipse = lambda *a, **k: empty
    skd.line = lambda *a, **k: empty
    sk.draw = skd
    sys.modules.setdefault("skimage", sk); sys.modules.setdefault("skimage.draw", skd)
    for m in ("matplotlib", "matplotlib.pyplot", "matplotlib.rcsetup", "matplotlib.patches"):
        if m not in sys.modules:
            mod = types.ModuleType(m)
            mod.__dict__.setdefault("use", lambda *a, **k: None)
            sys.modules[m] = mod
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].rcsetup = sys.modules["matplotlib.rcsetup"]
    import resource
    resource.setrlimit = lambda *a, **k: None
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self       # loss.py:37-38,102,109; utils.py:199
    if REF not in sys.path:
        sys.path.insert(0, REF)
    _installed = True


def load_reference_modules():
    """Returns (BDCN class, DenseNet2D class, utils module, helperfunctions module)."""
    install_shims()
    import bdcn_new
    import utils as ref_utils
    import helperfunctions as ref_hf
    from models.RITnet_v2 import DenseNet2D
    return bdcn_new.BDCN, DenseNet2D, ref_utils, ref_hf


def run_reference(setting, bdcn_sd, esf_sd, img, labels=None, cond=None, pupil_center=None,
                  elnorm=None):
    """calc_edge (utils.py:645-656) + DenseNet2D.__call__ (test.py:79-92) on CPU fp32."""
    BDCN, DenseNet2D, ref_utils, _ = load_reference_modules()
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        edge_model = BDCN()
        model = DenseNet2D(dict(setting))
    edge_model.load_state_dict(bdcn_sd)
    model.load_state_dict(esf_sd)
    edge_model.eval(); model.eval()
    B, _, H, W = img.shape
    if labels is None:                         # evaluate.py:118-120
        labels = torch.zeros((B, H, W), dtype=torch.long)
        labels[:, 0, 2] = 1; labels[:, 2, 2] = 2
    cond = torch.zeros((B, 4)) if cond is None else cond
    pupil_center = torch.zeros((B, 2)) if pupil_center is None else pupil_center
    elnorm = torch.zeros((B, 2, 5)) if elnorm is None else elnorm
    args = types.SimpleNamespace(prec=torch.float32, edge_thres=0)
    with torch.no_grad():
        edge = ref_utils.calc_edge(args, img, edge_model, torch.device("cpu"))
        # cond[:,1]=1 skips the per-sample seg losses on CPU but flips the iris-centre branch,
        # so keep cond as given and let the shimmed .cuda() carry the loss code.
        op, elPred, latent, loss, elOut = model(img, edge, labels, pupil_center, elnorm,
                                                torch.zeros((B, H, W)), torch.zeros((B, 3, H, W)),
                                                cond, torch.zeros(B, dtype=torch.long), 0)
        pred = ref_utils.get_predictions(op)
    return dict(edge=edge, op=op, elPred=elPred, latent=latent, elOut=elOut, loss=loss, pred=pred)
