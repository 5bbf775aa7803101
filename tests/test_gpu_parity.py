"""Parity of the CUDA path (through the C ABI / Python look-alikes) against the CPU oracle and the
reference-generated golden fixtures.  Tolerances are BASELINE.json's north_star bars:
>= 99.9 % per-pixel argmax agreement, mIoU within 0.1 pt, centres within 0.25 px, ellipse
parameters within 1e-2 relative (floor 1e-2 for near-zero values)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

CONFIGS = ["baseline", "baseline_edge", "baseline_adain", "baseline_adain_edge",
           "baseline_input_concat", "baseline_only_edge"]


@pytest.fixture(scope="module")
def env(golden_dir):
    import egn_b200
    from oracle import graph, synth
    assert torch.cuda.is_available(), "GPU tests need the B200 box"
    torch.set_num_threads(os.cpu_count() or 1)
    dev = torch.device("cuda:0")
    img = torch.from_numpy(np.load(os.path.join(golden_dir, "fwd_input.npz"))["img"])
    bsd = synth.make_bdcn_state(0)
    edge_model = egn_b200.BDCN()
    edge_model.load_state_dict(bsd)
    edge_model = edge_model.cuda().eval()
    edge_model.micro_batch = 4
    with torch.no_grad():
        edge_ref = graph.calc_edge(bsd, img)
    return dict(egn=egn_b200, graph=graph, synth=synth, dev=dev, img=img, bsd=bsd, edge_model=edge_model,
                edge_ref=edge_ref, golden=golden_dir)


def _model(env, cfg, mb=4):
    st = env["synth"].SETTINGS[cfg]
    esd = env["synth"].make_esf_state(st, 0)
    m = env["egn"].DenseNet2D(st)
    m.load_state_dict(esd)
    m = m.cuda().eval()
    m.micro_batch = mb
    return m, st, esd


def rel_err(a, b, floor=1e-2):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))


def test_native_library_is_loaded(env):
    import ctypes
    lib = env["egn"].load_library()
    assert isinstance(lib, ctypes.CDLL) and os.path.basename(lib._name) == "libegn.so"
    with open("/proc/self/maps") as f:
        assert "libegn.so" in f.read()


def test_bdcn_edge_parity(env):
    e = env["edge_model"].edge(env["img"].to(env["dev"]))
    e3 = env["edge_model"](torch.cat([env["img"]] * 3, 1).to(env["dev"]))[-1]
    gold = np.load(os.path.join(env["golden"], "fwd_baseline_edge.npz"))["edge"]
    assert e.shape == (2, 1, 240, 320) and e.dtype == torch.float32
    np.testing.assert_allclose(e.cpu().numpy(), env["edge_ref"].numpy(), atol=5e-4)
    np.testing.assert_allclose(e.cpu().numpy(), gold, atol=5e-4)          # the reference's own output
    np.testing.assert_allclose(e3.cpu().numpy(), gold, atol=5e-4)
    assert env["edge_model"].context().launch_count() > 0


def test_tensor_core_conv_matches_simt_companion(env):
    ctx = env["edge_model"].context(env["dev"])
    env["edge_model"].edge(env["img"].to(env["dev"]))
    for layer in ["features.conv1_2", "features.conv3_2", "features.conv4_1", "features.conv5_2",
                  "msblock1_1.conv", "msblock1_2.tail", "msblock3_3.tail", "msblock5_1.tail"]:
        d, r = ctx.conv_selfcheck(layer, 2)
        assert d <= 1e-4 * max(r, 1.0), (layer, d, r)


@pytest.mark.parametrize("cfg", CONFIGS)
def test_esf_forward_parity(env, cfg):
    m, st, esd = _model(env, cfg)
    dev, g = env["dev"], env["graph"]
    gold = np.load(os.path.join(env["golden"], f"fwd_{cfg}.npz"))
    with torch.no_grad():
        ref = g.esf_forward(esd, st, env["img"], env["edge_ref"])
        op, elPred, latent, loss, elOut = m(env["img"].to(dev), env["edge_ref"].to(dev), None, None, None, None, None,
                                            torch.zeros(2, 4, device=dev), 0, 0)
        pred = env["egn"].get_predictions(op, m)
        cond1 = torch.zeros(2, 4, device=dev); cond1[:, 1] = 1
        elPred_nomask = m(env["img"].to(dev), env["edge_ref"].to(dev), None, None, None, None, None, cond1, 0, 0)[1]
    assert op.shape == (2, 3, 240, 320) and elPred.shape == (2, 10) and latent.shape == (2, 153)
    assert loss.shape == (1,) and elOut.shape == (2, 10) and pred.dtype == torch.int64 and not pred.is_cuda
    pref = g.get_predictions(ref["op"]).numpy()
    assert (pred.numpy() == pref).mean() >= 0.999
    assert (pred.numpy() == gold["pred"]).mean() >= 0.999                 # vs the reference itself
    # centres: normalised -> pixels (0.5*W, 0.5*H scaling of utils.py:636-643)
    scale = np.array([160.0, 120.0])
    for sl in (slice(0, 2), slice(5, 7)):
        assert np.abs((elPred.cpu().numpy()[:, sl] - gold["elPred"][:, sl]) * scale).max() < 0.25
        assert np.abs((elOut.cpu().numpy()[:, sl] - gold["elOut"][:, sl]) * scale).max() < 0.25
    # ellipse parameters (axes, angle): 1e-2 relative; the centre entries are judged in pixels above
    par = [2, 3, 4, 7, 8, 9]
    assert rel_err(elOut.cpu().numpy()[:, par], gold["elOut"][:, par]) < 1e-2
    assert rel_err(elPred.cpu().numpy()[:, par], gold["elPred"][:, par]) < 1e-2
    assert rel_err(elPred_nomask.cpu().numpy()[:, par], gold["elPred_nomask"][:, par]) < 1e-2
    assert np.abs((elPred_nomask.cpu().numpy()[:, 0:2] - gold["elPred_nomask"][:, 0:2]) * scale).max() < 0.25
    assert rel_err(latent.cpu().numpy(), gold["latent"]) < 1e-2
    np.testing.assert_allclose(op.cpu().numpy()[:, :, ::4, ::4], gold["op_s4"], atol=5e-3)


def test_esf_tensor_core_layers_match_simt(env):
    m, st, esd = _model(env, "baseline_edge")
    dev = env["dev"]
    with torch.no_grad():
        m(env["img"].to(dev), env["edge_ref"].to(dev), None, None, None, None, None, torch.zeros(2, 4, device=dev), 0, 0)
    ctx = m.context(dev)
    for layer in ["enc.head.conv2", "enc.down_block1.conv21", "enc.down_block2.conv31", "enc.down_block3.conv1",
                  "enc.bottleneck.TD.conv", "dec.up_block4.conv21", "dec.up_block3.conv11", "dec.up_block1.conv22",
                  "dec.final.conv1", "dec.final.conv2", "elReg.c1"]:
        d, r = ctx.conv_selfcheck(layer, 2)
        assert d <= 1e-4 * max(r, 1.0), (layer, d, r)


def test_ragged_micro_batches_and_frame_independence(env):
    """B=5 with micro-batch 2 (ragged tail) must equal the same frames run alone; a frame's result
    must not depend on its neighbours (InstanceNorm / AdaIN are per-sample)."""
    m, st, esd = _model(env, "baseline_edge", mb=2)
    dev = env["dev"]
    em = env["egn"].BDCN(); em.load_state_dict(env["bsd"]); em = em.cuda().eval(); em.micro_batch = 2
    x = torch.cat([env["img"], env["synth"].randn_frames(3, seed=5)], 0).to(dev)
    with torch.no_grad():
        e = em.edge(x)
        lo, eo, la, am, ep = m.infer(x, e, None)
        perm = torch.tensor([4, 2, 0, 3, 1], device=dev)
        e2 = em.edge(x[perm])
        lo2, eo2, la2, am2, ep2 = m.infer(x[perm], e2, None)
    assert torch.equal(e[perm], e2)
    assert torch.equal(lo[perm], lo2) and torch.equal(am[perm], am2)
    assert torch.equal(eo[perm], eo2) and torch.equal(ep[perm], ep2) and torch.equal(la[perm], la2)


def test_seg_post_against_oracle(env):
    g, dev = env["graph"], env["dev"]
    gen = torch.Generator().manual_seed(3)
    logits = torch.randn(3, 3, 240, 320, generator=gen) * 2
    logits[:, :, 100, 100] = 1.0                       # exact three-way tie -> class 0
    logits[0, 1, 50, 60] = logits[0, 2, 50, 60] = 9.0  # two-way tie -> first index (1)
    el = torch.randn(3, 10, generator=gen)
    ctx = env["edge_model"].context(dev)
    am, ep = ctx.seg_post(logits.to(dev), el.to(dev), None)
    assert torch.equal(am.cpu().to(torch.int64), g.get_predictions(logits))
    pup, iri = g.seg2pt_exact(logits[:, 2]), g.seg2pt_exact(-logits[:, 0])
    want = torch.cat([iri, el[:, 2:5], pup, el[:, 7:10]], 1)
    np.testing.assert_allclose(ep.cpu().numpy(), want.numpy(), atol=2e-5)
    cond = torch.zeros(3, 4); cond[:, 1] = 1
    _, ep2 = ctx.seg_post(logits.to(dev), el.to(dev), cond.to(dev))
    np.testing.assert_allclose(ep2.cpu().numpy()[:, 0:2], el[:, 5:7].numpy(), atol=0)
    cond[1, 1] = 0                                       # one sample with a mask flips the whole batch
    _, ep3 = ctx.seg_post(logits.to(dev), el.to(dev), cond.to(dev))
    np.testing.assert_allclose(ep3.cpu().numpy(), want.numpy(), atol=2e-5)


def test_metrics_against_reference_fixture(env):
    g = np.load(os.path.join(env["golden"], "metrics.npz"))
    dev = env["dev"]
    ctx = env["edge_model"].context(dev)
    B = g["label"].shape[0]
    cond = torch.zeros(B, 4); cond[:, 1] = torch.from_numpy(g["cond"]); cond[:, 0] = cond[:, 1]
    for lab_dtype in (torch.uint8, torch.int64):
        acc = env["egn"].MetricAccumulator(dev)
        by = torch.empty(B, 3, device=dev)
        el = torch.zeros(B, 10)
        el[:, 5:7] = torch.from_numpy(g["ppred"])
        ctx.metrics_accumulate(torch.from_numpy(g["pred"]).to(dev), torch.from_numpy(g["label"]).to(dev).to(lab_dtype),
                               cond.to(dev), acc.acc, torch.from_numpy(g["ptrue"]).to(dev), torch.from_numpy(g["ptrue"]).to(dev),
                               el.to(dev), el.to(dev), by)
        r = acc.result()
        np.testing.assert_allclose(r["IoUs"], g["per"], rtol=1e-6)
        assert r["mIoU"] == pytest.approx(float(g["miou"]), rel=1e-6)
        np.testing.assert_allclose(by.cpu().numpy(), g["by"], rtol=1e-6, equal_nan=True)
        assert r["pupil_latent_px"] == pytest.approx(float(g["pd"]), rel=1e-5)
        assert r["frames"] == B


def test_ellipse_transform_and_refinement(env):
    g = np.load(os.path.join(env["golden"], "ellipse.npz"))
    dev = env["dev"]
    ctx = env["edge_model"].context(dev)
    n = len(g["params"]) // 2
    am = torch.zeros(n, 240, 320, dtype=torch.uint8, device=dev)
    out = ctx.ellipse_refine(am, torch.from_numpy(g["params"][:2 * n]).float().reshape(n, 2, 5), refine=False).cpu().numpy()
    np.testing.assert_allclose(out.reshape(-1, 5), g["transformed"][:2 * n, :5], rtol=2e-5, atol=2e-4)
    # refinement: masks come in (iris, pupil) pairs per synthetic eye
    masks = np.stack([np.unpackbits(m)[:240 * 320].reshape(240, 320) for m in g["masks"]])
    k = len(masks) // 2
    seg = np.zeros((k, 240, 320), np.uint8)
    Hinv = np.array([[2 / 320, 0, -1], [0, 2 / 240, -1], [0, 0, 1.0]])
    ell = np.zeros((k, 2, 5), np.float32)
    for i in range(k):
        seg[i][masks[2 * i].astype(bool)] = 1
        seg[i][masks[2 * i + 1].astype(bool)] = 2
        for w in range(2):
            ell[i, w] = env["graph"].ellipse_transform(g["inits"][2 * i + w], Hinv)[:5]
    # one degenerate frame: no pupil pixel at all (seg_count == 0 and, for the tiny start ellipse below, an
    # empty raster too -> IoU = 0/0 = NaN, every comparison of the descent is False, the start is returned)
    seg = np.concatenate([seg, (seg[:1] == 1).astype(np.uint8)], 0)
    tiny = ell[:1].copy()
    tiny[0, 1] = [0.9, 0.9, 1e-4, 1e-4, 0.3]
    ell = np.concatenate([ell, tiny], 0)
    out = ctx.ellipse_refine(torch.from_numpy(seg).to(dev), torch.from_numpy(ell), refine=True).cpu().numpy()
    deg = lambda p: [p[0], p[1], p[2], p[3], p[4] * 180 / 3.14159]
    worst = 0.0
    for i in range(k + 1):
        for w in range(2):
            m = seg[i] == (w + 1)
            got = out[i, w]
            # the oracle starts from exactly what the device starts from: the float32 normalised parameters
            # the ABI carries, mapped to pixels in float64 (evaluate.py:141-146)
            start = env["graph"].ellipse_norm_to_px(ell[i, w].astype(np.float64))
            want = env["graph"].refine_ellipse(m, start)
            np.testing.assert_allclose(got[:2], want[:2], rtol=1e-9, atol=1e-9)      # the search never moves the centre
            # a, b, theta: north_star's "ellipse parameters within 1e-2 relative" (floor 1e-2)
            err = rel_err(got[2:5], want[2:5])
            worst = max(worst, err)
            assert err < 1e-2, (i, w, got, want)
            if i < k:
                # and against the reference's own search from its float64 start (fixture): same objective value
                ref = g["refined"][2 * i + w]
                assert env["graph"].ell_iou(m, deg(got)) >= env["graph"].ell_iou(m, deg(ref)) - 0.01
                np.testing.assert_allclose(got[:2], ref[:2], atol=0.05)
    assert np.isfinite(out[:k]).all()
    # both schedules of the search (two directions per raster pass, and the reference's one candidate per pass)
    # walk the same path: same bits
    os.environ["EGN_REFINE_SPECULATE"] = "0"
    try:
        seq = ctx.ellipse_refine(torch.from_numpy(seg).to(dev), torch.from_numpy(ell), refine=True).cpu().numpy()
    finally:
        del os.environ["EGN_REFINE_SPECULATE"]
    np.testing.assert_array_equal(seq, out)
    # degenerate pupil: NaN objective, parameters come back unrefined
    np.testing.assert_allclose(out[k, 1], env["graph"].ellipse_norm_to_px(ell[k, 1].astype(np.float64)), rtol=1e-9)
    print("ellipse refinement: worst relative deviation of (a, b, theta) from the oracle %.2e" % worst)


def test_end_to_end_metrics_on_synthetic_eyes(env):
    """calc_acc-style evaluation (test.py:75-252) on labelled synthetic eyes: engine vs oracle."""
    egn, g, synth, dev = env["egn"], env["graph"], env["synth"], env["dev"]
    m, st, esd = _model(env, "baseline_edge")
    eb = synth.synthetic_eye_batch(100, 4)
    batch = tuple(torch.from_numpy(eb[k]) for k in ("img", "label", "spatW", "distMap", "pupil_center", "iris_center",
                                                   "elNorm", "cond", "imInfo"))
    acc = egn.MetricAccumulator(dev)
    with torch.no_grad():
        logits, el_pred, el_out, argmax = egn.evaluate_batch(m, env["edge_model"], batch, acc)
        e_ref = g.calc_edge(env["bsd"], batch[0])
        ref = g.esf_forward(esd, st, batch[0], e_ref)
    r = acc.result()
    pref = g.get_predictions(ref["op"]).numpy()
    miou_ref, per_ref, _ = g.seg_metrics(eb["label"], pref, eb["cond"][:, 1])
    assert abs(r["mIoU"] - miou_ref) * 100 < 0.1
    d_ref = g.point_metric(eb["pupil_center"], ref["elPred"][:, 5:7].numpy(), eb["cond"][:, 1], (240, 320))[0]
    assert abs(r["pupil_seg_px"] - d_ref) < 0.25
    d_ref = g.point_metric(eb["iris_center"], ref["elOut"][:, 0:2].numpy(), eb["cond"][:, 1], (240, 320))[0]
    assert abs(r["iris_latent_px"] - d_ref) < 0.25
    assert r["frames"] == 4


def test_evaluate_per_image_path(env):
    """evaluate.py:112-166 equivalent on a real frame pair."""
    egn, g, dev = env["egn"], env["graph"], env["dev"]
    m, st, esd = _model(env, "baseline_edge")
    fr = np.load(os.path.join(env["golden"], "frames_u8.npz"))["frames"][:1]
    x = egn.preprocess_frames_u8(fr, dev)
    np.testing.assert_allclose(x.cpu().numpy()[0, 0], g.preprocess_frame_u8(fr[0]), atol=1e-5)
    edge_map, seg_map, pupil, iris = egn.evaluate_ellseg_on_image(x, m, env["edge_model"])
    assert edge_map.shape == (240, 320) and seg_map.shape == (240, 320) and pupil.shape == (5,) and iris.shape == (5,)
    with torch.no_grad():
        e_ref = g.calc_edge(env["bsd"], x.cpu())
        ref = g.esf_forward(esd, st, x.cpu(), e_ref)
    pref = g.get_predictions(ref["op"]).numpy()[0]
    assert (seg_map == pref).mean() >= 0.999
    want = g.refine_ellipse(pref == 2, g.ellipse_norm_to_px(ref["elPred"][0, 5:10].numpy()))
    got_iou = g.ell_iou(seg_map == 2, [pupil[0], pupil[1], pupil[2], pupil[3], pupil[4] * 180 / 3.14159])
    want_iou = g.ell_iou(pref == 2, [want[0], want[1], want[2], want[3], want[4] * 180 / 3.14159])
    assert got_iou >= want_iou - 0.02 or not np.isfinite(want_iou)
    # the refinement itself, pinned on the engine's own segmentation and float32 elPred: the oracle's search
    # from the same start and the same mask must return the same (a, b, theta) for both ellipses
    with torch.no_grad():
        lo, eo, la, am, ep = m.infer(x, env["edge_model"].edge(x), None)
        ell = m.context(dev).ellipse_refine(am, ep, True).cpu().numpy()
    ep, seg2 = ep.cpu().numpy().astype(np.float64), am.cpu().numpy()[0]
    for got, cls, sl in ((ell[0, 0], 1, slice(0, 5)), (ell[0, 1], 2, slice(5, 10))):
        start = g.ellipse_norm_to_px(ep[0, sl])
        w = g.refine_ellipse(seg2 == cls, start)
        np.testing.assert_allclose(got[:2], w[:2], rtol=1e-9, atol=1e-9)
        assert rel_err(got[2:5], w[2:5]) < 1e-2, (cls, got, w)


def test_preprocess_u8_matches_numpy_zscore(env):
    """evaluate.py:102-103: (img - mean) / std in float64 -> float32, on every committed real frame."""
    egn, g, dev = env["egn"], env["graph"], env["dev"]
    fr = np.load(os.path.join(env["golden"], "frames_u8.npz"))["frames"]
    x = egn.preprocess_frames_u8(fr, dev).cpu().numpy()
    want = np.stack([g.preprocess_frame_u8(f) for f in fr])[:, None]
    assert x.shape == want.shape and x.dtype == np.float32
    np.testing.assert_allclose(x, want, rtol=0, atol=5e-7)
    flat = np.full((1, 240, 320), 7, np.uint8); flat[0, 0, 0] = 9      # near-constant frame: tiny std, no NaN
    y = egn.preprocess_frames_u8(flat, dev).cpu().numpy()
    np.testing.assert_allclose(y[0, 0], g.preprocess_frame_u8(flat[0]), rtol=1e-6, atol=1e-6)


def test_real_frames_batch_parity(env):
    """Eight real eye crops (videos/example1.avi via tests/golden/frames_u8.npz) through the u8 ingest,
    BDCN and ESF-Net with a ragged micro-batch; every frame must meet the argmax / centre bars."""
    egn, g, dev = env["egn"], env["graph"], env["dev"]
    m, st, esd = _model(env, "baseline_edge", mb=3)
    fr = np.load(os.path.join(env["golden"], "frames_u8.npz"))["frames"][:8]
    x = egn.preprocess_frames_u8(fr, dev)
    em = env["egn"].BDCN(); em.load_state_dict(env["bsd"]); em = em.cuda().eval(); em.micro_batch = 3
    with torch.no_grad():
        e = em.edge(x)
        lo, eo, la, am, ep = m.infer(x, e, None)
        e_ref = g.calc_edge(env["bsd"], x.cpu())
        ref = g.esf_forward(esd, st, x.cpu(), e_ref)
    np.testing.assert_allclose(e.cpu().numpy(), e_ref.numpy(), atol=5e-4)
    agree = (am.cpu().to(torch.int64) == g.get_predictions(ref["op"])).float().flatten(1).mean(1)
    assert agree.min().item() >= 0.999, agree
    scale = np.array([160.0, 120.0])
    for sl in (slice(0, 2), slice(5, 7)):
        assert np.abs((ep.cpu().numpy()[:, sl] - ref["elPred"].numpy()[:, sl]) * scale).max() < 0.25
    par = [2, 3, 4, 7, 8, 9]
    assert rel_err(eo.cpu().numpy()[:, par], ref["elOut"].numpy()[:, par]) < 1e-2


def test_throughput_tiling_parity(env, monkeypatch):
    """The small micro-batches of these tests plan the convolutions in latency mode (more, smaller
    CTA tiles); bench.py's micro-batch of 64 frames uses the throughput tiling.  EGN_TC_NO_LATENCY_MODE
    pins that tiling at a small batch so the exact kernel configuration that is benchmarked meets the
    same parity bars, end to end (BDCN edge -> ESF-Net -> argmax / centres / ellipse parameters)."""
    monkeypatch.setenv("EGN_TC_NO_LATENCY_MODE", "1")
    egn, g, dev = env["egn"], env["graph"], env["dev"]
    em = egn.BDCN(); em.load_state_dict(env["bsd"]); em = em.cuda().eval(); em.micro_batch = 2
    m, st, esd = _model(env, "baseline_edge", mb=2)
    gold = np.load(os.path.join(env["golden"], "fwd_baseline_edge.npz"))
    with torch.no_grad():
        e = em.edge(env["img"].to(dev))
        lo, eo, la, am, ep = m.infer(env["img"].to(dev), e, None)
    np.testing.assert_allclose(e.cpu().numpy(), gold["edge"], atol=5e-4)
    assert (am.cpu().numpy() == gold["pred"]).mean() >= 0.999
    scale = np.array([160.0, 120.0])
    for sl in (slice(0, 2), slice(5, 7)):
        assert np.abs((ep.cpu().numpy()[:, sl] - gold["elPred"][:, sl]) * scale).max() < 0.25
    par = [2, 3, 4, 7, 8, 9]
    assert rel_err(eo.cpu().numpy()[:, par], gold["elOut"][:, par]) < 1e-2
    for layer in ["features.conv4_2", "features.conv5_1", "msblock4_1.conv"]:
        d, r = em.context(dev).conv_selfcheck(layer, 2)
        assert d <= 1e-4 * max(r, 1.0), (layer, d, r)


def test_calc_acc_mirror_on_synthetic_loader(env, capsys):
    """egn_b200.calc_acc (the test.calc_acc mirror, test.py:32-252) over a two-batch loader with
    unequal batch sizes against the oracle's per-batch metrics aggregated the reference's way."""
    egn, g, synth, dev = env["egn"], env["graph"], env["synth"], env["dev"]
    m, st, esd = _model(env, "baseline_edge")
    keys = ("img", "label", "spatW", "distMap", "pupil_center", "iris_center", "elNorm", "cond", "imInfo")
    loader = []
    for start, n in ((200, 3), (300, 2)):
        eb = synth.synthetic_eye_batch(start, n)
        loader.append(tuple(torch.from_numpy(eb[k]) for k in keys))
    by = []
    out = egn.calc_acc(None, loader, m, env["edge_model"], dev, return_all=True, iou_by_sample_out=by)
    printed = capsys.readouterr().out
    assert "mIoU:" in printed and "Segmentation PUPIL dist. Mean:" in printed
    ious_b, dpl, dil, dps, dis = [], [], [], [], []
    with torch.no_grad():
        for batch in loader:
            e_ref = g.calc_edge(env["bsd"], batch[0])
            ref = g.esf_forward(esd, st, batch[0], e_ref)
            pref = g.get_predictions(ref["op"]).numpy()
            c = batch[7].numpy()
            ious_b.append(g.seg_metrics(batch[1].numpy(), pref, c[:, 1])[1])
            dpl.append(g.point_metric(batch[4].numpy(), ref["elOut"][:, 5:7].numpy(), c[:, 0], (240, 320))[0])
            dil.append(g.point_metric(batch[5].numpy(), ref["elOut"][:, 0:2].numpy(), c[:, 1], (240, 320))[0])
            dps.append(g.point_metric(batch[4].numpy(), ref["elPred"][:, 5:7].numpy(), c[:, 1], (240, 320))[0])
            dis.append(g.point_metric(batch[5].numpy(), ref["elPred"][:, 0:2].numpy(), c[:, 1], (240, 320))[0])
    ious_ref = np.nanmean(np.stack(ious_b), 0)
    assert np.abs(out[0] - ious_ref).max() * 100 < 0.1                      # mIoU within 0.1 pt, per class
    for got, want in zip(out[1:], (dpl, dil, dps, dis)):
        assert abs(got - np.nanmean(want)) < 0.25                            # centres within 0.25 px
    assert len(by) == 2 and by[0].shape == (3, 3) and by[1].shape == (2, 3)


def test_forward_loss_slot(env):
    """The loss slot of DenseNet2D.forward (get_allLoss, RITnet_v2.py:312-323,372-440) on the device:
    (1) egn_forward_loss on the fixture's inputs against the reference's own totals, for both label
    dtypes and for a batch without masks; (2) through the module call, against the oracle's loss on
    the oracle's forward outputs."""
    egn, g, synth, dev = env["egn"], env["graph"], env["synth"], env["dev"]
    gold = np.load(os.path.join(env["golden"], "loss.npz"))
    eb = synth.synthetic_eye_batch(500, 4)
    cond = torch.from_numpy(eb["cond"]).clone().float(); cond[2, 1] = 1
    op = synth.smooth_logits(eb["label"], seed=5)
    sw, dm = synth.loss_maps(eb["label"], seed=6)
    tgt, pc, en = torch.from_numpy(eb["label"]), torch.from_numpy(eb["pupil_center"]).float(), torch.from_numpy(eb["elNorm"]).float()
    el_out = torch.from_numpy(gold["el_out"])
    ctx = env["edge_model"].context(dev)
    for cnd, keys in ((cond, (("total_a0", 0.0), ("total_a5", 0.5), ("total_a10", 1.0))),
                      (torch.tensor([[0, 1, 0, 0]] * 4).float(), (("total_nomask", 0.5),))):
        am, el_pred = ctx.seg_post(op.to(dev).contiguous(), el_out.to(dev), cnd.to(dev))
        for key, alpha in keys:
            for lab in (tgt.to(torch.uint8), tgt.to(torch.int64)):
                loss = ctx.forward_loss(op.to(dev).contiguous(), lab, sw, dm, cnd, pc, en, el_out.to(dev), el_pred, alpha)
                assert loss.shape == (1,)
                assert float(loss.item()) == pytest.approx(float(gold[key]), rel=2e-4), (key, lab.dtype)
    # through the module: the 4th output of the reference's 5-tuple
    m, st, esd = _model(env, "baseline_edge")
    x = torch.from_numpy(eb["img"])
    with torch.no_grad():
        e_ref = g.calc_edge(env["bsd"], x)
        ref = g.esf_forward(esd, st, x, e_ref)
        want, _ = g.all_loss(ref["op"], ref["elOut"], tgt.long(), pc, en, sw, dm, cond, 0.5)
        out = m(x.to(dev), e_ref.to(dev), tgt.long().to(dev), pc.to(dev), en.to(dev), sw.to(dev), dm.to(dev),
                cond.to(dev), torch.zeros(4, dtype=torch.long, device=dev), 0.5)
    assert out[3].shape == (1,) and out[3].is_cuda
    assert float(out[3].item()) == pytest.approx(float(want), rel=5e-3)


def test_config0_baseline_batch16_calc_acc(env, capsys):
    """BASELINE.json configs[0]: configs/baseline.yaml (no edge input to the net), batch 16, through the
    calc_acc path - the engine against the oracle on labelled synthetic eyes, at BASELINE's bars."""
    egn, g, synth, dev = env["egn"], env["graph"], env["synth"], env["dev"]
    m, st, esd = _model(env, "baseline", mb=16)
    keys = ("img", "label", "spatW", "distMap", "pupil_center", "iris_center", "elNorm", "cond", "imInfo")
    eb = synth.synthetic_eye_batch(700, 16)
    batch = tuple(torch.from_numpy(eb[k]) for k in keys)
    em = egn.BDCN(); em.load_state_dict(env["bsd"]); em = em.cuda().eval(); em.micro_batch = 16
    out = egn.calc_acc(None, [batch], m, em, dev, return_all=True)
    capsys.readouterr()
    with torch.no_grad():
        ref = g.esf_forward(esd, st, batch[0], None)
    pref = g.get_predictions(ref["op"]).numpy()
    c = eb["cond"]
    ious_ref = g.seg_metrics(eb["label"], pref, c[:, 1])[1]
    assert np.abs(out[0] - ious_ref).max() * 100 < 0.1
    want = (g.point_metric(eb["pupil_center"], ref["elOut"][:, 5:7].numpy(), c[:, 0], (240, 320))[0],
            g.point_metric(eb["iris_center"], ref["elOut"][:, 0:2].numpy(), c[:, 1], (240, 320))[0],
            g.point_metric(eb["pupil_center"], ref["elPred"][:, 5:7].numpy(), c[:, 1], (240, 320))[0],
            g.point_metric(eb["iris_center"], ref["elPred"][:, 0:2].numpy(), c[:, 1], (240, 320))[0])
    for got, w in zip(out[1:], want):
        assert abs(got - w) < 0.25


def test_batch_of_one_shapes_and_int_id(env):
    """evaluate.py:115-131 calls the model with B == 1, ID == 0 (an int) and zero-filled targets; the
    reference squeezes / unsqueezes around that case (loss.py:42, RITnet_v2.py:409-412) and still returns
    [1,3,H,W], [1,10], [1,153], [1], [1,10]."""
    egn, dev = env["egn"], env["dev"]
    m, st, esd = _model(env, "baseline_edge", mb=1)
    x = env["img"][:1].to(dev)
    e = env["edge_ref"][:1].to(dev)
    labels = torch.zeros((1, 240, 320), device=dev)
    labels[..., 0, 2] = 1; labels[..., 2, 2] = 2
    with torch.no_grad():
        op, elPred, latent, loss, elOut = m(x, e, labels.long(), torch.zeros((1, 2), device=dev),
                                            torch.zeros((1, 2, 5), device=dev), torch.zeros((1, 240, 320), device=dev),
                                            torch.zeros((1, 3, 240, 320), device=dev), torch.zeros((1, 4), device=dev), 0, 0)
    assert op.shape == (1, 3, 240, 320) and elPred.shape == (1, 10) and latent.shape == (1, 153)
    assert loss.shape == (1,) and elOut.shape == (1, 10) and torch.isfinite(loss).all()
    gold = np.load(os.path.join(env["golden"], "fwd_baseline_edge.npz"))
    assert (egn.get_predictions(op, m).numpy()[0] == gold["pred"][0]).mean() >= 0.999


def test_plain_c_client_matches_python_mirrors(env, tmp_path):
    """tools/abi_client.c (gcc, no torch, device memory from the CUDA runtime) drives the same C ABI with
    weight blobs on disk; its argmax / elPred / edge files must equal the Python mirrors' outputs bit for bit."""
    import subprocess
    from egn_b200.pack import pack_state_dict
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tools", "abi_client")
    if not os.path.isfile(exe):
        pytest.skip("tools/abi_client not built (run __graft_entry__.build())")
    egn, dev = env["egn"], env["dev"]
    m, st, esd = _model(env, "baseline_edge", mb=2)
    em = egn.BDCN(); em.load_state_dict(env["bsd"]); em = em.cuda().eval(); em.micro_batch = 2
    (tmp_path / "bdcn.blob").write_bytes(pack_state_dict(env["bsd"]))
    (tmp_path / "esf.blob").write_bytes(pack_state_dict(esd))
    env["img"].numpy().astype("<f4").tofile(str(tmp_path / "x.f32"))
    r = subprocess.run([exe, str(tmp_path / "bdcn.blob"), str(tmp_path / "esf.blob"), str(tmp_path / "x.f32"), "2",
                        str(tmp_path / "out")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    with torch.no_grad():
        e = em.edge(env["img"].to(dev))
        lo, eo, la, am, ep = m.infer(env["img"].to(dev), e, None)
    got_am = np.fromfile(str(tmp_path / "out.argmax.u8"), dtype=np.uint8).reshape(2, 240, 320)
    got_ep = np.fromfile(str(tmp_path / "out.elpred.f32"), dtype="<f4").reshape(2, 10)
    got_e = np.fromfile(str(tmp_path / "out.edge.f32"), dtype="<f4").reshape(2, 1, 240, 320)
    assert np.array_equal(got_e, e.cpu().numpy())
    assert np.array_equal(got_am, am.cpu().numpy())
    assert np.array_equal(got_ep, ep.cpu().numpy())


def test_full_size_batch_properties(env):
    """BASELINE.json's batch of 256 frames (too large for the CPU oracle) through size-independent
    properties: frames are independent, so a frame must produce the same segmentation / centres /
    ellipse parameters wherever it sits in the batch and in whichever micro-batch it lands, and the
    metric accumulators of the whole batch must equal the sum of those of its halves."""
    egn, synth, dev = env["egn"], env["synth"], env["dev"]
    m, st, esd = _model(env, "baseline_edge", mb=64)
    em = egn.BDCN(); em.load_state_dict(env["bsd"]); em = em.cuda().eval(); em.micro_batch = 64
    half = synth.randn_frames(128, seed=7)
    x = torch.cat([half, half.flip(0)], 0).to(dev)               # frame i == frame 255 - i
    lab = synth.evaluate_style_labels(256).to(torch.uint8).to(dev)
    cond = torch.zeros(256, 4, device=dev)
    c = torch.full((256, 2), 100.0, device=dev)
    with torch.no_grad():
        e = em.edge(x)
        lo, eo, la, am, ep = m.infer(x, e, cond)
    assert torch.isfinite(lo).all() and torch.isfinite(eo).all()
    assert (am[:128] == am[128:].flip(0)).float().mean().item() >= 0.99999
    np.testing.assert_allclose(ep[:128].cpu().numpy(), ep[128:].flip(0).cpu().numpy(), atol=1e-4)
    np.testing.assert_allclose(eo[:128].cpu().numpy(), eo[128:].flip(0).cpu().numpy(), atol=1e-4)
    np.testing.assert_allclose(e[:128].cpu().numpy(), e[128:].flip(0).cpu().numpy(), atol=1e-6)
    ctx = m.context(dev)
    whole, a, b = egn.MetricAccumulator(dev), egn.MetricAccumulator(dev), egn.MetricAccumulator(dev)
    ctx.metrics_accumulate(am, lab, cond, whole.acc, c, c, eo, ep)
    ctx.metrics_accumulate(am[:128].contiguous(), lab[:128].contiguous(), cond[:128], a.acc, c[:128], c[:128], eo[:128].contiguous(), ep[:128].contiguous())
    ctx.metrics_accumulate(am[128:].contiguous(), lab[128:].contiguous(), cond[128:], b.acc, c[128:], c[128:], eo[128:].contiguous(), ep[128:].contiguous())
    np.testing.assert_allclose(whole.acc.cpu().numpy(), (a.acc + b.acc).cpu().numpy(), rtol=1e-12)
    assert whole.result()["frames"] == 256


def test_bdcn_side_outputs_are_real_tensors(env):
    """BDCN.forward returns the reference's 11-entry list (bdcn_new.py:178-191): [-1] eagerly, the ten
    per-scale sigmoids on first access of any other entry - compared here with the oracle's full forward."""
    g, dev = env["graph"], env["dev"]
    x3 = torch.cat([env["img"]] * 3, 1)
    out = env["edge_model"](x3.to(dev))
    assert len(out) == 11 and not out._filled
    fuse = out[-1]
    assert not out._filled                              # what utils.calc_edge does: no extra work
    with torch.no_grad():
        ref = g.bdcn_forward(env["bsd"], x3, return_all=True)
    assert len(ref) == 11
    for i, (a, b) in enumerate(zip(out, ref)):          # iterating materialises the side outputs
        assert a.shape == (2, 1, 240, 320) and a.dtype == torch.float32 and a.is_cuda
        np.testing.assert_allclose(a.cpu().numpy(), b.numpy(), atol=5e-4, err_msg="output %d" % i)
    assert out._filled and out[10] is fuse
    assert torch.equal(env["edge_model"](x3.to(dev))[3], out[3])


def test_get_predictions_keyed_on_the_logits_tensor(env):
    m, st, esd = _model(env, "baseline_edge")
    dev, egn = env["dev"], env["egn"]
    with torch.no_grad():
        a = m(env["img"].to(dev), env["edge_ref"].to(dev), None, None, None, None, None, torch.zeros(2, 4, device=dev), 0, 0)[0]
        pa = egn.get_predictions(a, m)
        held = a.clone()
        b = m(env["img"].flip(0).to(dev), env["edge_ref"].flip(0).to(dev), None, None, None, None, None,
              torch.zeros(2, 4, device=dev), 0, 0)[0]
        pb = egn.get_predictions(b, m)
        assert torch.equal(pb, pa.flip(0))
        assert torch.equal(egn.get_predictions(held, m), pa)       # logits of the EARLIER call: not the cached argmax
        assert torch.equal(egn.get_predictions(held), pa)


def test_engine_info_reports_the_precision_policy(env):
    info = env["edge_model"].context(env["dev"]).info()
    assert info["products_per_mac"] == 3 and info["tensor_core_path"] == 1 and info["micro_batch"] == 4
    assert info["workspace_bytes"] > 0 and info["num_sms"] >= 100


def test_merged_launches_phase_lattice_tails_and_fused_pools(env):
    """Round-2 restructurings of the BDCN graph, each against the SIMT companion on the same buffers:
    msblock.conv riding with the VGG convolution that reads the same map (bdcn_new.py:120-160 / vgg16_c.py:65-88),
    the stage-1/2 MSBlock tails on the 4x4 polyphase lattice (bdcn_new.py:49-55), and pools 1-3 written by the
    producing epilogue (vgg16_c.py:15,20,27; the pooled maps feed conv2_1 / conv3_1 / conv4_1, so test_bdcn_edge_parity
    pins them end to end - here the launch list shows that pools 1-3 no longer run as kernels)."""
    em, dev = env["edge_model"], env["dev"]
    ctx = em.context(dev)
    em.edge(env["img"].to(dev))
    for layer in ["features.conv2_2", "msblock2_1.conv", "features.conv3_3", "msblock3_2.conv", "msblock1_1.tail",
                  "msblock1_2.tail", "msblock2_1.tail", "msblock2_2.tail"]:
        d, r = ctx.conv_selfcheck(layer, 2)
        assert d <= 1e-4 * max(r, 1.0), (layer, d, r)
    # no standalone launch is left for the merged blocks
    table = None
    ctx.profile(True)
    em.edge(env["img"].to(dev))
    torch.cuda.synchronize()
    table = ctx.profile_table()
    ctx.profile(False)
    names = [ln.split(",")[1] for ln in table.splitlines()[1:] if ln.startswith("conv,")]
    assert "features.conv1_2+msblock1_1.conv" in names and "msblock1_1.conv" not in names and "msblock2_1.conv" not in names
    assert not any(ln.startswith("aux,bdcn.maxpool") and int(ln.split(",")[9]) > 1 for ln in table.splitlines()), table


def test_workspace_is_liveness_shared(env):
    """One arena per net with interval packing (engine.cuh commit_acts): repeated forwards stay bit-identical
    although buffers share memory, and the footprint is well below one allocation per activation."""
    # micro-batches above 16 frames share buffers by liveness (streaming micro-batches keep dedicated buffers: their
    # independent branches run on a second stream, engine.cuh side_of)
    m, st, esd = _model(env, "baseline_edge", mb=24)
    dev = env["dev"]
    x, e = env["img"].to(dev), env["edge_ref"].to(dev)
    with torch.no_grad():
        a = m(x, e, None, None, None, None, None, torch.zeros(2, 4, device=dev), 0, 0)
        b = m(x, e, None, None, None, None, None, torch.zeros(2, 4, device=dev), 0, 0)
    assert torch.equal(a[0], b[0]) and torch.equal(a[4], b[4])
    info = m.context(dev).info()
    assert info["activation_bytes_unshared"] > 0
    assert info["workspace_bytes"] < 0.75 * info["activation_bytes_unshared"], info
    em = env["egn"].BDCN(); em.load_state_dict(env["bsd"]); em = em.cuda().eval(); em.micro_batch = 24
    e1, e2 = em.edge(x), em.edge(x)
    assert torch.equal(e1, e2)
    np.testing.assert_allclose(e1.cpu().numpy(), env["edge_ref"].numpy(), atol=5e-4)
    binfo = em.context(dev).info()
    assert binfo["workspace_bytes"] < 0.75 * binfo["activation_bytes_unshared"], binfo
    # the same frames through a streaming micro-batch (side-stream branches, dedicated buffers) give the same bits
    m2, _, _ = _model(env, "baseline_edge", mb=2)
    with torch.no_grad():
        c = m2(x, e, None, None, None, None, None, torch.zeros(2, 4, device=dev), 0, 0)
    assert m2.context(dev).info()["activation_bytes_unshared"] == 0
    np.testing.assert_allclose(c[4].cpu().numpy(), a[4].cpu().numpy(), atol=2e-5)
    egn = env["egn"]
    assert (egn.get_predictions(c[0], m2) == egn.get_predictions(a[0], m)).float().mean().item() > 0.9999


def test_shared_workspace_pool_between_the_two_modules(env):
    """egn_share_workspace: the BDCN context lends its (dead) buffers to the ESF-Net - alternating forwards of the two
    modules through ONE pool give the bits of separate arenas."""
    egn, dev = env["egn"], env["dev"]
    x, e_ref = env["img"].to(dev), env["edge_ref"].to(dev)
    m0, st, esd = _model(env, "baseline_edge", mb=24)
    with torch.no_grad():
        ref = m0(x, e_ref, None, None, None, None, None, torch.zeros(2, 4, device=dev), 0, 0)
    egn.share_workspace(True)
    try:
        em = egn.BDCN(); em.load_state_dict(env["bsd"]); em = em.cuda().eval(); em.micro_batch = 24
        m, _, _ = _model(env, "baseline_edge", mb=24)
        outs = []
        with torch.no_grad():
            for _ in range(2):
                edge = em.edge(x)
                outs.append((edge, m(x, e_ref, None, None, None, None, None, torch.zeros(2, 4, device=dev), 0, 0)))
    finally:
        egn.share_workspace(False)
    for edge, o in outs:
        np.testing.assert_allclose(edge.cpu().numpy(), env["edge_ref"].numpy(), atol=5e-4)
        assert torch.equal(o[0], ref[0]) and torch.equal(o[4], ref[4])
    bi, mi = em.context(dev).info(), m.context(dev).info()
    assert bi["shared_pool_bytes"] > 0 and bi["shared_pool_bytes"] == mi["shared_pool_bytes"]
    assert m0.context(dev).info()["shared_pool_bytes"] == 0


def test_api_leaves_the_callers_device_alone(env):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    egn, dev1 = env["egn"], torch.device("cuda:1")
    torch.cuda.set_device(0)
    em = egn.BDCN(); em.load_state_dict(env["bsd"]); em = em.to(dev1).eval(); em.micro_batch = 2
    m, st, esd = _model(env, "baseline_edge", mb=2)
    m = m.to(dev1)
    with torch.no_grad():
        e = em.edge(env["img"].to(dev1))
        lo, eo, la, am, ep = m.infer(env["img"].to(dev1), e, None)
    assert torch.cuda.current_device() == 0                        # two contexts on cuda:1, caller still on cuda:0
    gold = np.load(os.path.join(env["golden"], "fwd_baseline_edge.npz"))
    np.testing.assert_allclose(e.cpu().numpy(), gold["edge"], atol=5e-4)
    assert (am.cpu().numpy() == gold["pred"]).mean() >= 0.999
    e0 = env["edge_model"].edge(env["img"].to(env["dev"]))          # and the cuda:0 context still works
    np.testing.assert_allclose(e0.cpu().numpy(), gold["edge"], atol=5e-4)


def test_reference_data_parallel_wrap(env):
    """test.py:264-269,294-297 on a multi-GPU box: model = torch.nn.DataParallel(model); model.to(device).
    BDCN stays on cuda:0 (test.py:284); the wrapped ESF-Net scatters the batch over the GPUs."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    egn, g, dev = env["egn"], env["graph"], env["dev"]
    st = env["synth"].SETTINGS["baseline_edge"]
    esd = env["synth"].make_esf_state(st, 0)
    model = egn.DenseNet2D(st)
    model.load_state_dict(esd)
    model.micro_batch = 2
    model = torch.nn.DataParallel(model)
    model = model.to(dev).to(torch.float32)
    model.eval()
    x = torch.cat([env["img"], env["img"].flip(0)], 0)
    with torch.no_grad():
        edge = egn.calc_edge(None, x, env["edge_model"], dev)
        for _ in range(2):                                          # second call reuses the per-device contexts
            op, elPred, latent, loss, elOut = model(x.to(dev), edge, None, None, None, None, None,
                                                    torch.zeros(4, 4, device=dev), None, 0)
        pred = egn.get_predictions(op, model)
    assert op.shape == (4, 3, 240, 320) and op.device == dev and elOut.shape == (4, 10)
    prim = model.module
    assert set(prim._ctxs) >= {torch.device("cuda", 0), torch.device("cuda", 1)}
    gold = np.load(os.path.join(env["golden"], "fwd_baseline_edge.npz"))
    want = np.concatenate([gold["pred"], gold["pred"][::-1]], 0)
    assert (pred.numpy() == want).mean() >= 0.999
    scale = np.array([160.0, 120.0])
    wel = np.concatenate([gold["elPred"], gold["elPred"][::-1]], 0)
    for sl in (slice(0, 2), slice(5, 7)):
        assert np.abs((elPred.cpu().numpy()[:, sl] - wel[:, sl]) * scale).max() < 0.25
