"""Thin Python handle on an egn_ctx (include/egn.h): owns the context, passes raw device pointers
and the current torch CUDA stream.  PyTorch is used for device memory and streams only."""
import ctypes
import os

import numpy as np
import torch

from . import _lib
from .pack import pack_state_dict

H, W = 240, 320
NET_BDCN, NET_ESF = 0, 1
DEFAULT_MICRO_BATCH = int(os.environ.get("EGN_MICRO_BATCH", "16"))


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _f32c(t, device):
    return t.to(device=device, dtype=torch.float32).contiguous()


class Context:
    def __init__(self, device, setting=None, micro_batch=None):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.EgnError("egn_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        device = torch.device(device)
        if device.type != "cuda":
            raise _lib.EgnError("egn_b200 modules live on CUDA devices only (got %s)" % device)
        self.device = torch.device("cuda", device.index if device.index is not None else torch.cuda.current_device())
        s = setting or {}
        cfg = _lib.EgnConfig(int(s.get("add_edge", 0)), int(s.get("add_seg", 0)), int(s.get("seg_detach", 0)),
                             int(s.get("input_concat", 0)), int(s.get("only_edge", 0)), int(s.get("style_dim", 8)))
        h = ctypes.c_void_p()
        _lib.check(self.lib.egn_create(self.device.index, ctypes.byref(cfg), ctypes.byref(h)))
        self.h = h
        self.micro_batch = int(micro_batch or DEFAULT_MICRO_BATCH)
        _lib.check(self.lib.egn_plan(self.h, self.micro_batch))

    def close(self):
        if getattr(self, "h", None):
            self.lib.egn_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_weights(self, net, state_dict):
        blob = pack_state_dict(state_dict)
        buf = ctypes.create_string_buffer(blob, len(blob))
        _lib.check(self.lib.egn_set_weights(self.h, net, ctypes.cast(buf, ctypes.c_void_p), len(blob)))

    # -- forward entry points ------------------------------------------------------------------
    def bdcn_forward(self, x):
        """x: [B,1,H,W] grey or [B,3,H,W]; returns edge [B,1,H,W] fp32 on the device."""
        assert x.dim() == 4 and x.shape[2] == H and x.shape[3] == W and x.shape[1] in (1, 3), x.shape
        x = _f32c(x, self.device)
        out = torch.empty((x.shape[0], 1, H, W), dtype=torch.float32, device=self.device)
        _lib.check(self.lib.egn_bdcn_forward(self.h, _ptr(x), int(x.shape[1]), _ptr(out), int(x.shape[0]),
                                             _stream(self.device)))
        return out

    def bdcn_forward_all(self, x):
        """Full BDCN.forward return (bdcn_new.py:178-191): (edge [B,1,H,W], sides [10,B,1,H,W])."""
        assert x.dim() == 4 and x.shape[2] == H and x.shape[3] == W and x.shape[1] in (1, 3), x.shape
        x = _f32c(x, self.device)
        B = int(x.shape[0])
        out = torch.empty((B, 1, H, W), dtype=torch.float32, device=self.device)
        sides = torch.empty((10, B, 1, H, W), dtype=torch.float32, device=self.device)
        _lib.check(self.lib.egn_bdcn_forward_all(self.h, _ptr(x), int(x.shape[1]), _ptr(out), _ptr(sides), B,
                                                 _stream(self.device)))
        return out, sides

    def info(self):
        """What this context runs (include/egn.h egn_info_t) as a dict."""
        i = _lib.EgnInfo()
        _lib.check(self.lib.egn_info(self.h, ctypes.byref(i)))
        return {k: getattr(i, k) for k, _ in _lib.EgnInfo._fields_}

    def esf_forward(self, x, edge):
        assert x.dim() == 4 and tuple(x.shape[1:]) == (1, H, W), x.shape
        B = int(x.shape[0])
        x = _f32c(x, self.device)
        edge = _f32c(edge, self.device) if edge is not None else None
        logits = torch.empty((B, 3, H, W), dtype=torch.float32, device=self.device)
        el_out = torch.empty((B, 10), dtype=torch.float32, device=self.device)
        latent = torch.empty((B, 153), dtype=torch.float32, device=self.device)
        _lib.check(self.lib.egn_esf_forward(self.h, _ptr(x), _ptr(edge), _ptr(logits), _ptr(el_out), _ptr(latent), B,
                                            _stream(self.device)))
        return logits, el_out, latent

    def seg_post(self, logits, el_out, cond=None):
        B = int(logits.shape[0])
        argmax = torch.empty((B, H, W), dtype=torch.uint8, device=self.device)
        el_pred = torch.empty((B, 10), dtype=torch.float32, device=self.device)
        cond = _f32c(cond, self.device) if cond is not None else None
        _lib.check(self.lib.egn_seg_post(self.h, _ptr(logits), _ptr(el_out), _ptr(cond), _ptr(argmax), _ptr(el_pred), B,
                                         _stream(self.device)))
        return argmax, el_pred

    def metrics_accumulate(self, argmax, labels, cond, acc, pupil_c=None, iris_c=None, el_out=None, el_pred=None,
                           iou_by_sample=None):
        B = int(argmax.shape[0])
        assert labels.dtype in (torch.uint8, torch.int64) and labels.is_cuda and labels.is_contiguous()
        assert acc.dtype == torch.float64 and acc.numel() >= 16 and acc.is_cuda
        cond = _f32c(cond, self.device)
        pc = _f32c(pupil_c, self.device) if pupil_c is not None else None
        ic = _f32c(iris_c, self.device) if iris_c is not None else None
        _lib.check(self.lib.egn_metrics_accumulate(self.h, _ptr(argmax), _ptr(labels), int(labels.dtype == torch.int64),
                                                   _ptr(cond), _ptr(pc), _ptr(ic), _ptr(el_out), _ptr(el_pred), _ptr(acc),
                                                   _ptr(iou_by_sample), B, _stream(self.device)))

    def forward_loss(self, logits, target, spat_w, dist_map, cond, pupil_c, el_norm, el_out, el_pred, alpha=0.0):
        """get_allLoss (models/RITnet_v2.py:372-440) forward value: fp32 tensor [1] on the device."""
        B = int(logits.shape[0])
        tgt = target.to(self.device)
        if tgt.dtype not in (torch.uint8, torch.int64):
            tgt = tgt.to(torch.int64)
        tgt = tgt.contiguous()
        loss = torch.empty(1, dtype=torch.float32, device=self.device)
        sw, dm, cd = _f32c(spat_w, self.device), _f32c(dist_map, self.device), _f32c(cond, self.device)
        pc, en = _f32c(pupil_c, self.device), _f32c(el_norm.reshape(B, 10), self.device)
        assert tuple(sw.shape) == (B, H, W) and tuple(dm.shape) == (B, 3, H, W) and tuple(tgt.shape) == (B, H, W)
        _lib.check(self.lib.egn_forward_loss(self.h, _ptr(logits), _ptr(tgt), int(tgt.dtype == torch.int64), _ptr(sw),
                                             _ptr(dm), _ptr(cd), _ptr(pc), _ptr(en), _ptr(el_out), _ptr(el_pred),
                                             ctypes.c_float(float(alpha)), _ptr(loss), B, _stream(self.device)))
        return loss

    def ellipse_refine(self, argmax, ell_norm, refine=True):
        B = int(argmax.shape[0])
        ell = _f32c(ell_norm.reshape(B, 2, 5), self.device)
        out = torch.empty((B, 2, 5), dtype=torch.float64, device=self.device)
        _lib.check(self.lib.egn_ellipse_refine(self.h, _ptr(argmax), _ptr(ell), _ptr(out), int(bool(refine)), B,
                                               _stream(self.device)))
        return out

    def preprocess_u8(self, frames_u8):
        """frames_u8: [B,H,W] uint8 on the device -> z-scored fp32 [B,1,H,W] (evaluate.py:102-103)."""
        assert frames_u8.dtype == torch.uint8 and frames_u8.dim() == 3 and tuple(frames_u8.shape[1:]) == (H, W)
        f = frames_u8.to(self.device).contiguous()
        out = torch.empty((f.shape[0], 1, H, W), dtype=torch.float32, device=self.device)
        _lib.check(self.lib.egn_preprocess_u8(self.h, _ptr(f), _ptr(out), int(f.shape[0]), _stream(self.device)))
        return out

    # -- introspection -------------------------------------------------------------------------
    def launch_count(self):
        return int(self.lib.egn_launch_count(self.h))

    def flops_per_frame(self, net):
        return float(self.lib.egn_flops_per_frame(self.h, net))

    def profile(self, enable):
        _lib.check(self.lib.egn_profile(self.h, int(bool(enable))))

    def profile_read(self, reset=True):
        ms, fl, n = ctypes.c_double(), ctypes.c_double(), ctypes.c_longlong()
        _lib.check(self.lib.egn_profile_read(self.h, ctypes.byref(ms), ctypes.byref(fl), ctypes.byref(n), int(reset)))
        return ms.value, fl.value, n.value

    def profile_table(self):
        n = self.lib.egn_profile_table(self.h, None, 0)
        buf = ctypes.create_string_buffer(int(n))
        self.lib.egn_profile_table(self.h, buf, n)
        return buf.value.decode()

    def debug_read(self, name, frames):
        dims = (ctypes.c_int * 3)()
        n = self.lib.egn_debug_read(self.h, name.encode(), None, 0, frames, dims)
        if n < 0:
            raise _lib.EgnError(self.lib.egn_last_error().decode())
        out = np.empty(n, dtype=np.float32)
        n2 = self.lib.egn_debug_read(self.h, name.encode(), out.ctypes.data_as(ctypes.c_void_p), n, frames, dims)
        assert n2 == n
        return out.reshape(-1, dims[0], dims[1], dims[2])

    def conv_selfcheck(self, layer, frames=1):
        d, r = ctypes.c_double(), ctypes.c_double()
        _lib.check(self.lib.egn_conv_selfcheck(self.h, layer.encode(), frames, ctypes.byref(d), ctypes.byref(r)))
        return d.value, r.value
