"""Drop-in for the reference's ``bdcn_new.BDCN`` (bdcn_new.py:65-217) backed by libegn.so.

Same constructor, ``load_state_dict`` key set (182 tensors), ``.cuda()/.eval()``
and ``__call__(x[B,3,H,W]) -> list`` whose last entry is the fused edge map - the only output the
reference consumes (utils.py:649, evaluate.py:106).  The ten per-scale sigmoids are dead at
inference: they are computed lazily, on first access of any other list entry (``_BdcnOutputs``)."""
import torch

from ._modules import EngineBound
from .engine import NET_BDCN
from .shapes import bdcn_param_shapes


class BDCN(EngineBound):
    _net = NET_BDCN

    def __init__(self, pretrain=None, logger=None, rate=4):
        super().__init__()
        if rate != 4:
            raise ValueError("egn_b200.BDCN supports the reference default rate=4 only")
        self.pretrain = pretrain
        self.declare(bdcn_param_shapes(), seed=11)
        self._init_binding()
        if pretrain:
            sd = torch.load(pretrain, map_location="cpu")
            own = self.state_dict()
            own.update({k: v for k, v in sd.items() if k in own and "features" in k})
            self.load_state_dict(own)

    def edge(self, x):
        """x: [B,1,H,W] grey or [B,3,H,W] -> fused edge map [B,1,H,W] (= reference forward(x)[-1])."""
        if not x.is_cuda:
            raise RuntimeError("egn_b200.BDCN runs on CUDA tensors only (no CPU fallback)")
        return self._ensure_ctx(x.device).bdcn_forward(x)

    def forward(self, x):
        return _BdcnOutputs(self, x, self.edge(x))


class _BdcnOutputs(list):
    """The 11-entry list of bdcn_new.py:178-191.  ``[-1]`` (the fused map, the only entry utils.calc_edge and
    evaluate.py read) is computed eagerly; the ten per-scale sigmoids are materialised on first access of
    any other entry (one more pass of the engine with the side outputs switched on), so a caller that
    iterates or indexes the list gets real ``[B,1,H,W]`` tensors, never placeholders."""

    def __init__(self, module, x, fuse):
        super().__init__([None] * 10 + [fuse])
        self._module, self._x, self._filled = module, x, False

    def _fill(self):
        if not self._filled:
            _, sides = self._module._ensure_ctx(self._x.device).bdcn_forward_all(self._x)
            for i in range(10):
                list.__setitem__(self, i, sides[i])
            self._filled, self._x = True, None

    def __getitem__(self, i):
        if not (isinstance(i, int) and i in (-1, 10)):
            self._fill()
        return super().__getitem__(i)

    def __iter__(self):
        self._fill()
        return super().__iter__()
