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
    change (load_state_dict / in-place edits bump tensor versions) or the module moves.

    Contexts are cached PER DEVICE on the module that owns the parameters.  ``nn.DataParallel``
    (test.py:264-269,296 wraps the model whenever more than one GPU is visible) replicates the module on
    every forward; a replica never owns a context - it resolves to its primary's cache entry for the
    device it was scattered to, which is created once from the primary's tensors and reused by later
    forwards.  A module only ever closes contexts it created itself."""
    _net = None
    _setting = None

    def _init_binding(self):
        self._ctxs = {}            # device -> (Context, key)
        self._items = None
        self._primary = None       # set on DataParallel replicas
        self.micro_batch = None

    def _replicate_for_data_parallel(self):
        replica = super()._replicate_for_data_parallel()
        replica._primary = self._primary if self._primary is not None else self
        replica._ctxs = None       # replicas own nothing
        replica._items = None
        return replica

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
        owner = self._primary if self._primary is not None else self
        device = torch.device(device)
        if device.type == "cuda" and device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        items = owner._tensors()
        key = (owner.micro_batch, tuple([(t._version, t.data_ptr()) for _, t in items]))
        hit = owner._ctxs.get(device)
        if hit is None or hit[1] != key:
            if hit is not None:
                hit[0].close()
            ctx = Context(device, owner._setting, owner.micro_batch)
            ctx.set_weights(owner._net, {k: t for k, t in items})
            owner._ctxs[device] = (ctx, key)
            return ctx
        return hit[0]

    def context(self, device=None):
        if device is None:
            owner = self._primary if self._primary is not None else self
            device = next(owner.parameters()).device
        return self._ensure_ctx(torch.device(device))
