"""Host-side mirrors of the reference's helper functions around the two modules
(utils.calc_edge, utils.get_predictions, the metric part of test.calc_acc, evaluate.py's per-image
path), each routed to libegn.so kernels."""
import numpy as np
import torch

H, W = 240, 320


def calc_edge(args, img, edge_model, device):
    """utils.py:645-656.  The reference builds cat(img,img,img); the engine reads the grey frame
    directly (conv1_1 weights summed over the three identical input channels)."""
    with torch.no_grad():
        edge = edge_model.edge(img.to(device))
    if getattr(args, "edge_thres", 0) == 1:
        edge = torch.where(edge >= 0.1, torch.ones_like(edge), edge)
    return edge


def get_predictions(output, model=None):
    """utils.py:65-81: argmax over the class dimension as an int64 CPU tensor [B,H,W].

    With ``model`` (the DenseNet2D that produced exactly this ``output`` tensor - same storage, same
    version, same shape) the device-side u8 argmax of that forward is reused, so only B*H*W bytes cross
    PCIe instead of the fp32 logits; any other logits tensor goes through the argmax kernel again."""
    if model is not None and getattr(model, "last_argmax", None) is not None \
            and getattr(model, "_last_logits_key", None) == (output.data_ptr(), output._version, tuple(output.shape)):
        return model.last_argmax.cpu().to(torch.int64)
    if output.is_cuda:
        from .engine import Context
        ctx = _scratch_ctx(output.device)
        el = torch.zeros((output.shape[0], 10), dtype=torch.float32, device=output.device)
        am, _ = ctx.seg_post(output.to(torch.float32).contiguous(), el)
        return am.cpu().to(torch.int64)
    raise RuntimeError("egn_b200.get_predictions needs CUDA logits (no CPU fallback)")


_scratch = {}


def _scratch_ctx(device):
    from .engine import Context
    key = str(device)
    if key not in _scratch:
        _scratch[key] = Context(device, None, 1)
    return _scratch[key]


def preprocess_frames_u8(frames_u8, device):
    """evaluate.py:102-103 / CurriculumLib.py:139-140: per-frame z-score of uint8 frames
    [B,H,W] -> fp32 [B,1,H,W] on the device (population std in float64, like numpy); only the
    uint8 frames cross PCIe."""
    f = torch.as_tensor(frames_u8)
    if f.dtype != torch.uint8:
        raise TypeError("preprocess_frames_u8 expects uint8 frames")
    return _scratch_ctx(torch.device(device)).preprocess_u8(f.to(device))


class MetricAccumulator:
    """Device-side accumulators of test.calc_acc's metrics (test.py:159-252), layout in
    csrc/post.cuh.  ``all_reduce`` sums them across ranks (the path's only collective)."""
    N = 16

    def __init__(self, device):
        self.acc = torch.zeros(self.N, dtype=torch.float64, device=device)

    def all_reduce(self):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.acc, op=dist.ReduceOp.SUM)
        return self

    @staticmethod
    def summarize(acc):
        a = np.asarray(acc, dtype=np.float64)
        with np.errstate(all="ignore"):
            per = np.where(a[3:6] > 0, a[0:3] / a[3:6], np.nan)
            d = np.where(a[10:14] > 0, a[6:10] / a[10:14], np.nan)
        return {"mIoU": float(np.nanmean(per)) if np.isfinite(per).any() else float("nan"),
                "IoUs": per, "pupil_latent_px": float(d[0]), "iris_latent_px": float(d[1]),
                "pupil_seg_px": float(d[2]), "iris_seg_px": float(d[3]), "frames": int(a[14])}

    def result(self):
        return self.summarize(self.acc.cpu().numpy())


def evaluate_batch(model, edge_model, batch, accumulator, args=None):
    """One iteration of test.calc_acc's loop (test.py:75-214) on the engine: edge, forward,
    argmax/centres and metric accumulation, all on the device.  ``batch`` is the reference
    DataLoader's 9-tuple (CurriculumLib.py:139-166)."""
    img, labels, spatW, distMap, pupil_center, iris_center, elNorm, cond, imInfo = batch
    dev = next(model.parameters()).device
    edge = calc_edge(args, img, edge_model, dev) if model.setting["add_edge"] or model.setting["input_concat"] \
        or model.setting["only_edge"] or (args is not None and getattr(args, "always_edge", True)) else None
    cond_d = cond.to(dev, torch.float32)
    logits, el_out, latent, argmax, el_pred = model.infer(img.to(dev), edge, cond_d)
    lab = labels.to(dev)
    if lab.dtype not in (torch.uint8, torch.int64):
        lab = lab.to(torch.int64)
    model.context(dev).metrics_accumulate(argmax, lab.contiguous(), cond_d, accumulator.acc,
                                          pupil_center.to(dev), iris_center.to(dev), el_out, el_pred)
    return logits, el_pred, el_out, argmax


def summarize_batches(acc_rows):
    """Reduces per-batch accumulator rows [nbatches,16] the way test.calc_acc does (test.py:215-252):
    per batch the nanmean over samples of each class IoU / each centre distance (what getSeg_metrics
    and getPoint_metric return, utils.py:120-162), then nanmean over batches, then the mean of the three
    class IoUs.  Returns (ious[3], pupil_latent, iris_latent, pupil_seg, iris_seg)."""
    a = np.asarray(acc_rows, dtype=np.float64).reshape(-1, 16)
    with np.errstate(all="ignore"):
        iou_b = np.where(a[:, 3:6] > 0, a[:, 0:3] / a[:, 3:6], np.nan)
        dist_b = np.where(a[:, 10:14] > 0, a[:, 6:10] / a[:, 10:14], np.nan)
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", category=RuntimeWarning)
            ious = np.nanmean(iou_b, axis=0)
            d = np.nanmean(dist_b, axis=0)
    return ious, float(d[0]), float(d[1]), float(d[2]), float(d[3])


def calc_acc(args, testloader, model, edge_model, device, return_all=False, disp=0, iou_by_sample_out=None):
    """Drop-in for test.calc_acc (test.py:32-252) on the engine: same loop over the DataLoader's
    9-tuples, same prints, same return values - but edge, forward, argmax, centres and the per-sample
    Jaccard / centre-distance metrics all stay on the device; one small D2H of the per-batch
    accumulators happens after the last batch.  ``iou_by_sample_out`` (a list) receives the per-sample
    IoUs [B,3] of every batch, the data test.py:219-230 pickles when args.record_iou == 1."""
    model.eval()
    dev = torch.device(device)
    rows, by_sample = [], []
    with torch.no_grad():
        for bt, batchdata in enumerate(testloader):
            if getattr(args, "test_normal", False) and bt > 10:
                break
            acc = MetricAccumulator(dev)
            img, labels, spatW, distMap, pupil_center, iris_center, elNorm, cond, imInfo = batchdata
            edge = calc_edge(args, img, edge_model, dev)
            cond_d = cond.to(dev, torch.float32)
            logits, el_out, latent, argmax, el_pred = model.infer(img.to(dev), edge, cond_d)
            lab = labels.to(dev)
            if lab.dtype not in (torch.uint8, torch.int64):
                lab = lab.to(torch.int64)
            by = torch.empty((img.shape[0], 3), dtype=torch.float32, device=dev) \
                if (iou_by_sample_out is not None or getattr(args, "record_iou", 0) == 1) else None
            model.context(dev).metrics_accumulate(argmax, lab.contiguous(), cond_d, acc.acc, pupil_center.to(dev),
                                                  iris_center.to(dev), el_out, el_pred, by)
            rows.append(acc.acc)
            if by is not None:
                by_sample.append(by)
    a = torch.stack(rows).cpu().numpy() if rows else np.zeros((0, 16))
    ious, d_pl, d_il, d_ps, d_is = summarize_batches(a)
    if iou_by_sample_out is not None:
        iou_by_sample_out.extend(b.cpu().numpy() for b in by_sample)
    print('mIoU: {}. IoUs: {}'.format(np.mean(ious), ious))
    print('Latent space PUPIL dist. Mean: {}'.format(d_pl))
    print('Segmentation PUPIL dist. Mean: {}'.format(d_ps))
    print('Latent space IRIS dist. Mean: {}'.format(d_il))
    print('Segmentation IRIS dist. Mean: {}'.format(d_is))
    if return_all:
        return ious, d_pl, d_il, d_ps, d_is
    return np.mean(ious), d_pl, d_il


def evaluate_ellseg_on_image(frame, model, edge_model, device=None, refine=True):
    """evaluate.py:112-166: edge, forward, argmax, normalised->pixel ellipses and the IoU
    refinement, returning (edge_map, seg_map, pupil_ellipse, iris_ellipse) as numpy arrays.
    frame: [B,1,H,W] (the reference is B == 1; any B works and adds a leading axis)."""
    assert frame.dim() == 4, "Frame must be [B,1,H,W]"
    dev = device or next(model.parameters()).device
    with torch.no_grad():
        fr = frame.to(dev)
        edge = edge_model.edge(fr)
        logits, el_out, latent, argmax, el_pred = model.infer(fr, edge, None)
        ell = model.context(dev).ellipse_refine(argmax, el_pred, refine)
    ell = ell.cpu().numpy()
    e, s = edge.squeeze(1).cpu().numpy(), argmax.cpu().numpy().astype(np.int64)
    if frame.shape[0] == 1:
        return e[0], s[0], ell[0, 1], ell[0, 0]
    return e, s, ell[:, 1], ell[:, 0]


def preprocess_frame(img, op_shape=(H, W), align_width=True):
    """evaluate.py:69-104: bring a grey frame [h,w] to ``op_shape`` (width-aligned Lanczos resize, then
    symmetric vertical zero-padding or cropping), z-score it and return ``(tensor[1,H,W] float32,
    scale_shift)``.  Same arithmetic as the reference (numpy float64 mean / population std); the
    reference's crop branch indexes with floats and raises on current numpy - here it crops the
    rows the reference intended.  Frames that are already ``op_shape`` uint8 can skip this and go
    through ``preprocess_frames_u8`` (z-score on the device)."""
    import cv2
    if not align_width:
        raise SystemExit('Height alignment not implemented! Exiting ...')
    img = np.asarray(img)
    if op_shape[1] != img.shape[1]:
        sc = op_shape[1] / img.shape[1]
        width, height = int(img.shape[1] * sc), int(img.shape[0] * sc)
        img = cv2.resize(img, (width, height), interpolation=cv2.INTER_LANCZOS4)
        if op_shape[0] > img.shape[0]:
            pad = op_shape[0] - img.shape[0]
            img = np.pad(img, ((pad // 2, pad - pad // 2), (0, 0)))
            scale_shift = (sc, pad)
        elif op_shape[0] < img.shape[0]:
            pad = op_shape[0] - img.shape[0]                      # negative: rows to drop
            top = (-pad) // 2
            img = img[top:top + op_shape[0], ...]
            scale_shift = (sc, pad)
        else:
            scale_shift = (sc, 0)
    else:
        scale_shift = (1, 0)
    img = (img - img.mean()) / img.std()
    return torch.from_numpy(np.ascontiguousarray(img)).unsqueeze(0).to(torch.float32), scale_shift


def rescale_to_original(edge_map, seg_map, pupil_ellipse, iris_ellipse, scale_shift, orig_shape):
    """evaluate.py:168-192: map the 240x320 outputs back to the original frame - ellipse centres / axes
    un-shifted and un-scaled (the angle is kept), maps re-padded (cropped input) or un-padded (padded
    input) and resized with nearest neighbour."""
    import cv2
    pupil_ellipse = np.array(pupil_ellipse, dtype=np.float64)
    iris_ellipse = np.array(iris_ellipse, dtype=np.float64)
    for e in (pupil_ellipse, iris_ellipse):
        e[1] = e[1] - np.floor(scale_shift[1] // 2)
        e[:-1] = e[:-1] * (1 / scale_shift[0])
    seg_map = np.asarray(seg_map)
    edge_map = np.asarray(edge_map)
    sh = scale_shift[1]
    if sh < 0:
        seg_map = np.pad(seg_map, ((-sh // 2, -sh // 2), (0, 0)))
        edge_map = np.pad(edge_map, ((-sh // 2, -sh // 2), (0, 0)))
    elif sh > 0:
        # the frame had been padded by sh rows: drop them again (the reference then calls np.pad with
        # negative widths, evaluate.py:185-188, which raises on every numpy - only the crop is kept)
        seg_map = seg_map[sh // 2:seg_map.shape[0] - (sh - sh // 2), ...]
        edge_map = edge_map[sh // 2:edge_map.shape[0] - (sh - sh // 2), ...]
    seg_map = cv2.resize(seg_map, (orig_shape[1], orig_shape[0]), interpolation=cv2.INTER_NEAREST)
    edge_map = cv2.resize(edge_map, (orig_shape[1], orig_shape[0]), interpolation=cv2.INTER_NEAREST)
    return edge_map, seg_map, pupil_ellipse, iris_ellipse


def shard_frames(total, rank, world):
    """Contiguous block partition of `total` frames over `world` ranks: returns
    (start, count); the first total % world ranks take one extra frame."""
    base, rem = divmod(int(total), int(world))
    start = rank * base + min(rank, rem)
    return start, base + (1 if rank < rem else 0)
