"""egn_b200 - B200-native engine behind the reference's own call surface for the edge+ESF-Net path.

    from egn_b200 import BDCN, DenseNet2D          # same ctor / load_state_dict / __call__
    import egn_b200; egn_b200.install()            # or: make `from bdcn_new import BDCN` and
                                                   # `from models.RITnet_v2 import DenseNet2D`
                                                   # resolve to these classes (test.py, evaluate.py)

The compute lives in libegn.so (hand-written sm_100a CUDA, C ABI in include/egn.h)."""
import sys
import types

from ._lib import EgnError, LIB_PATH, check, load as load_library   # noqa: F401
from .bdcn_new import BDCN                                            # noqa: F401
from .ritnet_v2 import DenseNet2D, getSizes                           # noqa: F401
from .engine import Context, NET_BDCN, NET_ESF                        # noqa: F401
from .hostapi import (calc_edge, get_predictions, evaluate_batch, MetricAccumulator,     # noqa: F401
                      preprocess_frames_u8, evaluate_ellseg_on_image, shard_frames, calc_acc,
                      summarize_batches, preprocess_frame, rescale_to_original)


def share_workspace(enable=True):
    """Opt-in (include/egn.h egn_share_workspace): every context created on a device from now on draws its activation
    arena from ONE pool, so the BDCN module's buffers are lent to the ESF-Net (37 instead of 56 GB at micro-batch 256).
    Only for callers that drive their modules in stream order on one stream (calc_acc, bench.py)."""
    check(load_library().egn_share_workspace(int(bool(enable))))


def install():
    """Registers module aliases so unmodified reference scripts import the look-alikes."""
    from . import bdcn_new as _b, ritnet_v2 as _r
    sys.modules["bdcn_new"] = _b
    pkg = sys.modules.get("models")
    if pkg is None:
        pkg = types.ModuleType("models")
        pkg.__path__ = []
        sys.modules["models"] = pkg
    sys.modules["models.RITnet_v2"] = _r
    pkg.RITnet_v2 = _r
