"""Shared machinery of the two nn.Module look-alikes: a parameter tree with the reference's key
names, and lazy binding of those parameters to an egn engine context."""
import math

import torch
import torch.nn as nn

from .engine import Context


class ParamTree(nn.Module):
    """Holds parameters/buffers under dotted names (nested anonymous sub-modules), so that
    state_dict()/load_state_dict(strict=True) behave like the reference module's."""

    def declare(self, shapes, seed=0):
        g = torch.Generator().manual_seed(seed)
        for name, shape in shapes.items():
            parts = name.split(".")
            mod = self
            for p in parts[:-1]:
                if p not in mod._modules:
                    mod.add_module(p, nn.Module())
                mod = mod._modules[p]
            leaf = parts[-1]
            if leaf == "num_batches_tracked":
                mod.register_buffer(leaf, torch.tensor(0, dtype=torch.long))
            elif leaf == "running_mean":
                mod.register_buffer(leaf, torch.zeros(shape))
            elif leaf == "running_var":
                mod.register_buffer(leaf, torch.ones(shape))
            elif len(shape) == 1:
                init = torch.ones(shape) if (leaf == "weight") else torch.zeros(shape)
                mod.register_parameter(leaf, nn.Parameter(init))
            else:
                fan_in = 1
                for s in shape[1:]:
                    fan_in *= s
                w = torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_in)
                mod.register_parameter(leaf, nn.Parameter(w))


class EngineBound(ParamTree):
    """Binds the module's current tensors to a device context on first use and re-binds when they
    change (load_state_dict / in-place edits bump tensor versions) or the module moves."""
    _net = None
    _setting = None

    def _init_binding(self):
        self._ctx = None
        self._bound_key = None
        self._items = None
        self.micro_batch = None

    def _apply(self, fn, *a, **k):
        # .to()/.cuda()/.float() may replace buffer tensors: drop the cached tensor list
        self._items = None
        return super()._apply(fn, *a, **k)

    def _tensors(self):
        # the tensor objects are stable between _apply calls (load_state_dict copies in place), so the
        # per-call staleness check is two attribute reads per tensor, not a state_dict() walk - this
        # sits on the streaming path (batch 1, evaluate.py:112-166)
        if self._items is None:
            self._items = list(self.state_dict(keep_vars=True).items())
        return self._items

    def _ensure_ctx(self, device):
        items = self._tensors()
        key = (device, self.micro_batch, tuple([(t._version, t.data_ptr()) for _, t in items]))
        if self._ctx is None or key != self._bound_key:
            if self._ctx is not None:
                self._ctx.close()
            ctx = Context(device, self._setting, self.micro_batch)
            ctx.set_weights(self._net, {k: t for k, t in items})
            self._ctx, self._bound_key = ctx, key
        return self._ctx

    def context(self, device=None):
        if device is None:
            device = next(self.parameters()).device
        return self._ensure_ctx(torch.device(device))
