"""Pins the BENCHMARKED configuration to the reference (VERDICT r1 "next" #1): 256 frames through ONE
256-frame micro-batch with the throughput tiling - exactly what bench.py times - with the reference's own
golden frames planted at batch positions 0 / 127 / 255 and eight real eye crops spread through the batch
(bench.plant).  The planted frames must meet all four north_star bars against
tests/golden/planted_<config>.npz (outputs of the unmodified reference, oracle/make_golden_planted.py) and the
golden pair additionally against tests/golden/fwd_<config>.npz and the oracle port."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _close(module):
    for ctx, _ in list(module._ctxs.values()):
        ctx.close()
    module._ctxs.clear()


@pytest.mark.parametrize("cfg", ["baseline_edge", "baseline_adain_edge"])
def test_benchmarked_configuration_meets_the_parity_bars(cfg, golden_dir):
    import bench
    import egn_b200
    from oracle import graph, synth
    dev = torch.device("cuda:0")
    B = 256
    st = synth.SETTINGS[cfg]
    bsd, esd = synth.make_bdcn_state(0), synth.make_esf_state(st, 0)
    em = egn_b200.BDCN(); em.load_state_dict(bsd); em = em.to(dev).eval()
    m = egn_b200.DenseNet2D(st); m.load_state_dict(esd); m = m.to(dev).eval()
    em.micro_batch = m.micro_batch = 256
    x = synth.randn_frames(B, seed=100)
    pos = bench.plant(x, B)
    assert [p for p, _ in pos[:3]] == [0, 127, 255] and len(pos) == 11
    try:
        with torch.no_grad():
            xd = x.to(dev)
            edge = egn_b200.calc_edge(None, xd, em, dev)
            op, el_pred, latent, _, el_out = m(xd, edge, None, None, None, None, None, torch.zeros(B, 4, device=dev), 0, 0)
            pred = egn_b200.get_predictions(op, m)
        info = m.context(dev).info()
        assert info["micro_batch"] == 256 and info["products_per_mac"] == 3 and info["tensor_core_path"] == 1
        par = bench.planted_parity(cfg, pos, pred.to(torch.uint8), el_pred, el_out)
        print(cfg, par)
        assert par["planted_frames"] == 11
        assert par["argmax_agreement"] >= 0.999, par
        assert par["centre_px"] < 0.25, par
        assert par["ell_rel"] < 1e-2, par
        # the golden pair against the original two-frame fixture (reference-generated) as well ...
        gold = np.load(os.path.join(golden_dir, "fwd_%s.npz" % cfg))
        for p, i in pos[:3]:
            assert (pred[p].numpy() == gold["pred"][i]).mean() >= 0.999
            np.testing.assert_allclose(op[p].cpu().numpy()[:, ::4, ::4], gold["op_s4"][i], atol=5e-3)
            if cfg == "baseline_edge":
                np.testing.assert_allclose(edge[p].cpu().numpy(), gold["edge"][i], atol=5e-4)
        # ... and two of the real crops against the oracle port run here (edge map and mIoU bars)
        sel = [pos[3][0], pos[10][0]]
        with torch.no_grad():
            e_ref = graph.calc_edge(bsd, x[sel])
            ref = graph.esf_forward(esd, st, x[sel], e_ref)
        np.testing.assert_allclose(edge[sel].cpu().numpy(), e_ref.numpy(), atol=5e-4)
        pref = graph.get_predictions(ref["op"]).numpy()
        lab = synth.evaluate_style_labels(2).numpy()
        for k in range(2):
            # mIoU of engine vs oracle prediction against the same labels: within 0.1 pt
            a = graph.seg_metrics(lab[k:k + 1], pred[sel[k]].numpy()[None], np.zeros(1))[0]
            b = graph.seg_metrics(lab[k:k + 1], pref[k:k + 1], np.zeros(1))[0]
            assert abs(a - b) * 100 < 0.1 or (np.isnan(a) and np.isnan(b))
        # no planted frame may depend on its position: frame 0 == frame 255 (same input) bit for bit
        assert torch.equal(pred[0], pred[255]) and torch.equal(el_out[0], el_out[255])
    finally:
        _close(em); _close(m)
        torch.cuda.empty_cache()
