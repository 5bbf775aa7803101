"""Drop-in for the reference's ``models.RITnet_v2.DenseNet2D`` (models/RITnet_v2.py:203-354).

Same constructor (yaml ``setting`` dict), state_dict key set, ``.to()/.cuda()/.eval()`` and
``__call__`` signature / 5-tuple return; the forward runs in libegn.so.  Always evaluates with the
BatchNorm running statistics (what test.py:49 does).  The loss slot of the reference's forward
(get_allLoss, RITnet_v2.py:312-323,372-440) is evaluated on the device (forward value, no autograd)
when the caller passes the target tensors, and is zeros(1) when they are None."""
import torch

from ._modules import EngineBound
from .engine import NET_ESF
from .shapes import esf_param_shapes, esf_sizes


def getSizes(chz, growth, blks=4):
    """models/RITnet_v2.py:15-29."""
    return esf_sizes(chz, growth, blks)


class DenseNet2D(EngineBound):
    _net = NET_ESF

    def __init__(self, setting, chz=32, growth=1.2, actfunc=None, norm=None, selfCorr=False, disentangle=False):
        super().__init__()
        if chz != 32 or abs(growth - 1.2) > 1e-12:
            raise ValueError("egn_b200.DenseNet2D supports the reference defaults chz=32, growth=1.2 only")
        self.sizes = getSizes(chz, growth)
        self.toggle = True
        self.selfCorr = selfCorr
        self.disentangle = disentangle
        self.disentangle_alpha = 2
        self.setting = dict(setting)
        self._setting = self.setting
        assert self.setting["input_concat"] + self.setting["add_edge"] < 2, "edge can use only 1 time!"
        if self.setting["add_edge"] == 1:
            assert self.setting["feature_channels"] * 2 == 306
        if selfCorr or disentangle:
            raise NotImplementedError("selfCorr / disentangle are training-only paths (out of scope)")
        self.declare(esf_param_shapes(self.setting), seed=12)
        self._init_binding()

    def setDatasetInfo(self, numSets=2):
        self.numSets = numSets      # training-only head (RITnet_v2.py:240-249); nothing to build

    def infer(self, x, x_edge=None, cond=None):
        """Inference core: logits, elOut, latent, argmax (u8), elPred - all device tensors."""
        if not x.is_cuda:
            raise RuntimeError("egn_b200.DenseNet2D runs on CUDA tensors only (no CPU fallback)")
        ctx = self._ensure_ctx(x.device)
        logits, el_out, latent = ctx.esf_forward(x, x_edge)
        argmax, el_pred = ctx.seg_post(logits, el_out, cond)
        return logits, el_out, latent, argmax, el_pred

    def forward(self, x, x_edge, target=None, pupil_center=None, elNorm=None, spatWts=None, distMap=None,
                cond=None, ID=None, alpha=0):
        logits, el_out, latent, argmax, el_pred = self.infer(x, x_edge, cond)
        # u8 [B,H,W] on device, reused by egn_b200.get_predictions for exactly this logits tensor
        self.last_argmax = argmax
        self._last_logits_key = (logits.data_ptr(), logits._version, tuple(logits.shape))
        tensors = (target, pupil_center, elNorm, spatWts, distMap, cond)
        if all(torch.is_tensor(t) for t in tensors):
            # the loss slot of the reference forward (RITnet_v2.py:312-323): forward value only
            loss = self._ensure_ctx(x.device).forward_loss(logits, target, spatWts, distMap, cond, pupil_center, elNorm,
                                                           el_out, el_pred, alpha)
        else:
            loss = torch.zeros(1, dtype=torch.float32, device=x.device)
        return logits, el_pred, latent, loss, el_out
