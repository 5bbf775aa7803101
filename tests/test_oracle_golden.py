"""Pins oracle/graph.py against fixtures produced by the unmodified reference
(oracle/make_golden.py, run in the build container)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import graph, synth

CONFIGS = ["baseline", "baseline_edge", "baseline_adain", "baseline_adain_edge",
           "baseline_input_concat", "baseline_only_edge"]


@pytest.fixture(scope="module")
def edge_and_img(golden_dir):
    img = torch.from_numpy(np.load(os.path.join(golden_dir, "fwd_input.npz"))["img"])
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    with torch.no_grad():
        edge = graph.calc_edge(synth.make_bdcn_state(0), img)
    return img, edge


def test_state_key_tables_match_reference(golden_dir):
    keys = json.load(open(os.path.join(golden_dir, "state_keys.json")))
    assert {k: list(v) for k, v in synth.bdcn_shapes().items()} == keys["bdcn"]
    for c in CONFIGS:
        mine = {k: list(v) for k, v in synth.esf_shapes(synth.SETTINGS[c]).items()}
        assert mine == keys[c], c


def test_bdcn_edge_matches_reference(edge_and_img, golden_dir):
    _, edge = edge_and_img
    ref = np.load(os.path.join(golden_dir, "fwd_baseline_edge.npz"))["edge"]
    assert edge.shape == (2, 1, 240, 320)
    np.testing.assert_allclose(edge.numpy(), ref, atol=2e-5, rtol=0)


@pytest.mark.parametrize("cfg", CONFIGS)
def test_esf_forward_matches_reference(cfg, edge_and_img, golden_dir):
    img, edge = edge_and_img
    g = np.load(os.path.join(golden_dir, f"fwd_{cfg}.npz"))
    st = synth.SETTINGS[cfg]
    sd = synth.make_esf_state(st, 0)
    with torch.no_grad():
        r = graph.esf_forward(sd, st, img, edge)
        r2 = graph.esf_forward(sd, st, img, edge, has_mask=torch.zeros(2, dtype=torch.bool))
    op = r["op"].numpy()
    np.testing.assert_allclose(op[:, :, ::4, ::4], g["op_s4"], atol=2e-4, rtol=1e-4)
    np.testing.assert_allclose(op.mean((2, 3)), g["op_mean"], atol=1e-4)
    pred = graph.get_predictions(r["op"]).numpy().astype(np.uint8)
    agree = (pred == g["pred"]).mean()
    # disagreements may only sit on near-ties of the reference logits
    assert agree > 0.9995, agree
    bad = pred != g["pred"]
    assert (g["margin_f16"].astype(np.float32)[bad] < 1e-3).all()
    np.testing.assert_allclose(r["elOut"].numpy(), g["elOut"], atol=2e-5)
    np.testing.assert_allclose(r["elPred"].numpy(), g["elPred"], atol=2e-5)
    np.testing.assert_allclose(r["latent"].numpy(), g["latent"], atol=2e-5)
    np.testing.assert_allclose(r2["elPred"].numpy(), g["elPred_nomask"], atol=2e-5)


def test_ellipse_transform_and_refine(golden_dir):
    g = np.load(os.path.join(golden_dir, "ellipse.npz"))
    Hm = np.array([[160.0, 0, 160.0], [0, 120.0, 120.0], [0, 0, 1.0]])
    for p, t in zip(g["params"], g["transformed"]):
        np.testing.assert_allclose(graph.ellipse_transform(p, Hm), t, rtol=1e-9, atol=1e-9)
    for mbits, init, ref, iou in zip(g["masks"], g["inits"], g["refined"], g["init_iou"]):
        m = np.unpackbits(mbits)[:240 * 320].reshape(240, 320).astype(bool)
        deg = np.array([init[0], init[1], init[2], init[3], init[4] * 180.0 / 3.14159])
        assert graph.ell_iou(m, deg) == pytest.approx(float(iou), abs=1e-7)
        np.testing.assert_allclose(graph.refine_ellipse(m, init), ref, rtol=0, atol=1e-9)


def test_metrics(golden_dir):
    g = np.load(os.path.join(golden_dir, "metrics.npz"))
    miou, per, by = graph.seg_metrics(g["label"], g["pred"], g["cond"])
    np.testing.assert_allclose(per, g["per"], rtol=1e-12)
    np.testing.assert_allclose(by, g["by"], rtol=1e-12, equal_nan=True)
    assert miou == pytest.approx(float(g["miou"]), rel=1e-12)
    pd, pds = graph.point_metric(g["ptrue"], g["ppred"], g["cond"], (240, 320))
    assert pd == pytest.approx(float(g["pd"]), rel=1e-6)
    np.testing.assert_allclose(pds, g["pds"], rtol=1e-5, atol=1e-5)


def _loss_inputs():
    from oracle import synth
    eb = synth.synthetic_eye_batch(500, 4)
    cond = torch.from_numpy(eb["cond"]).clone().float()
    cond[2, 1] = 1
    op = synth.smooth_logits(eb["label"], seed=5)
    sw, dm = synth.loss_maps(eb["label"], seed=6)
    return dict(op=op, tgt=torch.from_numpy(eb["label"]).long(), pc=torch.from_numpy(eb["pupil_center"]).float(),
                en=torch.from_numpy(eb["elNorm"]).float(), sw=sw, dm=dm, cond=cond)


def test_loss_slot_against_reference_fixture(golden_dir):
    """get_allLoss and its per-sample terms (models/RITnet_v2.py:372-440, loss.py:48-137): the oracle
    on regenerated seeded inputs against the values the real reference produced (oracle/make_golden_loss.py)."""
    from oracle import graph
    g = np.load(os.path.join(golden_dir, "loss.npz"))
    d = _loss_inputs()
    el_out = torch.from_numpy(g["el_out"])
    for key, alpha in (("total_a0", 0.0), ("total_a5", 0.5), ("total_a10", 1.0)):
        total, pcs = graph.all_loss(d["op"], el_out, d["tgt"], d["pc"], d["en"], d["sw"], d["dm"], d["cond"], alpha)
        assert float(total) == pytest.approx(float(g[key]), rel=1e-5)
        assert pcs.shape == (4, 2, 2)
    cn = d["cond"].clone(); cn[:, 1] = 1
    total, pcs = graph.all_loss(d["op"], el_out, d["tgt"], d["pc"], d["en"], d["sw"], d["dm"], cn, 0.5)
    assert float(total) == pytest.approx(float(g["total_nomask"]), rel=1e-5)
    np.testing.assert_allclose(pcs[:, 0].numpy(), el_out[:, 5:7].numpy())       # iris centre := elOut[:,5:7]
    for i in range(4):
        assert float(graph.surface_loss(d["op"][i], d["dm"][i])) == pytest.approx(float(g["surface"][i]), rel=1e-4, abs=1e-7)
        assert float(graph.wce_loss(d["op"][i], d["tgt"][i], d["sw"][i])) == pytest.approx(float(g["wce"][i]), rel=1e-5)
        assert float(graph.gdice_loss(d["op"][i], d["tgt"][i])) == pytest.approx(float(g["gdice"][i]), rel=1e-5)
