"""state_dict -> weight blob for egn_set_weights (format documented in csrc/engine.cuh::parse_blob).

All graph knowledge (concat layouts, tensor-core operand packing, BN folding, the collapsed BDCN
side chain) lives in the C++ engine; Python only hands over the reference's own tensors by name."""
import struct

import numpy as np
import torch

MAGIC = 0x574E4745  # 'EGNW'


def pack_state_dict(sd):
    parts = [struct.pack("<II", MAGIC, len(sd))]
    for name, t in sd.items():
        if name.startswith("module."):           # saved from nn.DataParallel (pytorchtools.py:113-123)
            name = name[len("module."):]
        a = t.detach().to("cpu", torch.float32).contiguous().numpy() if isinstance(t, torch.Tensor) \
            else np.asarray(t, dtype=np.float32)
        nb = name.encode("utf-8")
        parts.append(struct.pack("<I", len(nb)))
        parts.append(nb)
        parts.append(struct.pack("<I", a.ndim))
        parts.append(struct.pack("<%dq" % a.ndim, *a.shape))
        parts.append(np.ascontiguousarray(a, dtype="<f4").tobytes())
    return b"".join(parts)
