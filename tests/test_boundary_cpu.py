"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports everything
include/egn.h declares, the look-alike modules mirror the reference's state_dict key sets, the weight
blob format, and the multi-rank host logic (gloo, world_size 2).  No compute call touches a GPU."""
import json
import os
import re
import struct

import numpy as np
import pytest
import torch

import egn_b200
from egn_b200 import _lib
from egn_b200.pack import pack_state_dict, MAGIC
from egn_b200.shapes import bdcn_param_shapes, esf_param_shapes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CONFIGS = {
    "baseline": dict(add_seg=0, seg_detach=0, add_edge=0, feature_channels=153, style_dim=8, input_concat=0, only_edge=0),
    "baseline_edge": dict(add_seg=0, seg_detach=0, add_edge=1, feature_channels=153, style_dim=8, input_concat=0, only_edge=0),
    "baseline_adain": dict(add_seg=1, seg_detach=0, add_edge=0, feature_channels=153, style_dim=8, input_concat=0, only_edge=0),
    "baseline_adain_edge": dict(add_seg=1, seg_detach=0, add_edge=1, feature_channels=153, style_dim=8, input_concat=0, only_edge=0),
    "baseline_input_concat": dict(add_seg=0, seg_detach=0, add_edge=0, feature_channels=153, style_dim=8, input_concat=1, only_edge=0),
    "baseline_only_edge": dict(add_seg=0, seg_detach=0, add_edge=0, feature_channels=153, style_dim=8, input_concat=0, only_edge=1),
}


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "egn.h")).read()
    declared = set(re.findall(r"\b(egn_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(_lib.EXPORTS)
    assert lib.egn_version() >= 100


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    m = egn_b200.BDCN()
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 3, 240, 320))
    d = egn_b200.DenseNet2D(CONFIGS["baseline"])
    with pytest.raises(RuntimeError):
        d(torch.zeros(1, 1, 240, 320), torch.zeros(1, 1, 240, 320))
    with pytest.raises(_lib.EgnError):
        egn_b200.Context("cuda:0")


def test_state_dict_keys_match_reference(golden_dir):
    keys = json.load(open(os.path.join(golden_dir, "state_keys.json")))
    assert {k: list(v) for k, v in bdcn_param_shapes().items()} == keys["bdcn"]
    m = egn_b200.BDCN()
    assert {k: list(v.shape) for k, v in m.state_dict().items()} == keys["bdcn"]
    for name, st in CONFIGS.items():
        d = egn_b200.DenseNet2D(st)
        assert {k: list(v.shape) for k, v in d.state_dict().items()} == keys[name], name
        assert {k: list(v) for k, v in esf_param_shapes(st).items()} == keys[name], name


def test_checkpoint_containers_load_strict(tmp_path):
    """gen_00000016.pt {'a': sd} and baseline_edge_16.pkl {'state_dict': sd, 'epoch': n}."""
    from oracle import synth
    p1, p2 = tmp_path / "gen_00000016.pt", tmp_path / "baseline_edge_16.pkl"
    torch.save(synth.bdcn_checkpoint(0), p1)
    torch.save(synth.esf_checkpoint(synth.SETTINGS["baseline_edge"], 0), p2)
    m = egn_b200.BDCN()
    m.load_state_dict(torch.load(p1)["a"])
    d = egn_b200.DenseNet2D(CONFIGS["baseline_edge"])
    d.load_state_dict(torch.load(p2)["state_dict"])
    with pytest.raises(RuntimeError):
        d.load_state_dict({"bogus": torch.zeros(1)})
    with pytest.raises(AssertionError):
        egn_b200.DenseNet2D(dict(CONFIGS["baseline_edge"], input_concat=1))     # RITnet_v2.py:273


def test_blob_format_roundtrip():
    sd = {"a.weight": torch.arange(6, dtype=torch.float32).reshape(2, 3), "module.b": torch.tensor(3),
          "c": torch.zeros(0)}
    blob = pack_state_dict(sd)
    magic, count = struct.unpack_from("<II", blob, 0)
    assert magic == MAGIC and count == 3
    off = 8
    seen = {}
    for _ in range(count):
        (nl,) = struct.unpack_from("<I", blob, off); off += 4
        name = blob[off:off + nl].decode(); off += nl
        (nd,) = struct.unpack_from("<I", blob, off); off += 4
        dims = struct.unpack_from("<%dq" % nd, blob, off); off += 8 * nd
        n = int(np.prod(dims)) if nd else 1
        seen[name] = (dims, np.frombuffer(blob, "<f4", n, off).copy()); off += 4 * n
    assert off == len(blob)
    assert seen["a.weight"][0] == (2, 3) and seen["a.weight"][1].tolist() == [0, 1, 2, 3, 4, 5]
    assert "b" in seen and seen["b"][1].tolist() == [3.0]           # module. prefix stripped


def test_install_aliases():
    import sys
    saved = {k: sys.modules.get(k) for k in ("bdcn_new", "models", "models.RITnet_v2")}
    try:
        egn_b200.install()
        from bdcn_new import BDCN
        from models.RITnet_v2 import DenseNet2D
        assert BDCN is egn_b200.BDCN and DenseNet2D is egn_b200.DenseNet2D
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_shard_frames_partitions_exactly():
    for total in (0, 1, 7, 256, 4096, 4099):
        for world in (1, 2, 3, 8):
            spans = [egn_b200.shard_frames(total, r, world) for r in range(world)]
            assert sum(c for _, c in spans) == total
            pos = 0
            for s, c in spans:
                assert s == pos
                pos += c
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def test_metric_summary_matches_oracle_definition(golden_dir):
    from oracle import graph
    g = np.load(os.path.join(golden_dir, "metrics.npz"))
    _, per, by = graph.seg_metrics(g["label"], g["pred"], g["cond"])
    acc = np.zeros(16)
    for i in range(by.shape[0]):
        for c in range(3):
            if np.isfinite(by[i, c]):
                acc[c] += by[i, c]; acc[3 + c] += 1
    res = egn_b200.MetricAccumulator.summarize(acc)
    np.testing.assert_allclose(res["IoUs"], per, rtol=1e-12)
    assert res["mIoU"] == pytest.approx(float(np.nanmean(per)))


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    acc = egn_b200.MetricAccumulator("cpu")
    start, count = egn_b200.shard_frames(10, rank, world)
    # each rank "evaluates" its shard: IoU of frame i for class c is (i+1)/(10*(c+1))
    for i in range(start, start + count):
        for c in range(3):
            acc.acc[c] += (i + 1) / (10.0 * (c + 1)); acc.acc[3 + c] += 1
        acc.acc[6] += float(i); acc.acc[10] += 1; acc.acc[14] += 1
    acc.all_reduce()
    q.put((rank, acc.result()))
    dist.destroy_process_group()


def test_metric_all_reduce_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for _, r in res:
        assert r["frames"] == 10
        np.testing.assert_allclose(r["IoUs"], [0.55, 0.275, 0.55 / 3], rtol=1e-12)
        assert r["pupil_latent_px"] == pytest.approx(4.5)


def test_summarize_batches_matches_calc_acc_aggregation():
    """test.py:215-252 aggregates nanmean-over-batches of per-batch nanmeans (utils.py:120-162); the
    host reduction of the per-batch device accumulators must reproduce that, including batches where
    a class never occurs in the ground truth and batches without valid samples."""
    rng = np.random.default_rng(3)
    rows, iou_b, dist_b = [], [], []
    for b in range(5):
        n = int(rng.integers(1, 6))
        s = rng.random((n, 3))
        s[rng.random((n, 3)) < 0.3] = np.nan                 # class absent from GT -> NaN (getSeg_metrics)
        if b == 2:
            s[:] = np.nan                                     # no valid sample in this batch
        d = rng.random((n, 4)) * 10
        valid = rng.random((n, 4)) < 0.8
        a = np.zeros(16)
        a[0:3] = np.nansum(s, 0); a[3:6] = np.isfinite(s).sum(0)
        a[6:10] = (d * valid).sum(0); a[10:14] = valid.sum(0); a[14] = n
        rows.append(a)
        with np.errstate(all="ignore"):
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore", category=RuntimeWarning)
                iou_b.append(np.nanmean(s, 0))
            dist_b.append(np.where(valid.sum(0) > 0, (d * valid).sum(0) / np.maximum(valid.sum(0), 1), np.nan))
    ious, pl, il, ps, isg = egn_b200.summarize_batches(np.stack(rows))
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", category=RuntimeWarning)
        np.testing.assert_allclose(ious, np.nanmean(np.stack(iou_b), 0), rtol=1e-12)
        np.testing.assert_allclose([pl, il, ps, isg], np.nanmean(np.stack(dist_b), 0), rtol=1e-12)


def test_header_is_plain_c(tmp_path):
    """include/egn.h must be consumable by a C compiler (no C++ / torch types in the ABI), and the
    plain-C client of the whole path must compile against it."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "t.c"
    src.write_text('#include "egn.h"\nint main(void) { egn_config c = {0, 0, 0, 0, 0, 8}; (void)c; return egn_version() < 0; }\n')
    r = subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-fsyntax-only", "-I" + os.path.join(root, "include"), str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    if os.path.isdir("/usr/local/cuda/include"):
        r = subprocess.run(["gcc", "-std=c99", "-Wall", "-fsyntax-only", "-I" + os.path.join(root, "include"),
                            "-I/usr/local/cuda/include", os.path.join(root, "tools", "abi_client.c")], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr


def test_data_parallel_replicas_never_own_a_context():
    """test.py:264-269,296 wraps the model in nn.DataParallel on multi-GPU boxes; replicate() calls
    _replicate_for_data_parallel() on every forward.  A replica must resolve to its primary's per-device
    context cache and must not be able to close anything (shallow-copied __dict__ shared _ctx)."""
    d = egn_b200.DenseNet2D(CONFIGS["baseline_edge"])
    d.micro_batch = 3

    class FakeCtx:
        closed = False

        def close(self):
            self.closed = True
    sentinel = FakeCtx()
    d._ctxs[torch.device("cuda", 0)] = (sentinel, "key")
    r = d._replicate_for_data_parallel()
    assert r._primary is d and r._ctxs is None and r._items is None
    rr = r._replicate_for_data_parallel()           # a replica of a replica still points at the owner
    assert rr._primary is d
    assert d._ctxs[torch.device("cuda", 0)][0] is sentinel and not sentinel.closed
    # the reference's own wrap, literally (no forward without a GPU): state_dict keys gain the module. prefix
    # that pack_state_dict strips (pytorchtools.py:113-123)
    wrapped = torch.nn.DataParallel(d)
    assert all(k.startswith("module.") for k in wrapped.state_dict())
    b = egn_b200.BDCN()
    assert b._replicate_for_data_parallel()._primary is b


def test_get_predictions_never_reuses_a_stale_argmax():
    """The cached device argmax belongs to ONE logits tensor (storage, version, shape); any other tensor of
    the same batch size must not get it.  CPU-only: the fallback for CPU logits raises."""
    d = egn_b200.DenseNet2D(CONFIGS["baseline"])
    logits = torch.zeros(2, 3, 240, 320)
    d.last_argmax = torch.ones(2, 240, 320, dtype=torch.uint8)
    d._last_logits_key = (logits.data_ptr(), logits._version, tuple(logits.shape))
    assert egn_b200.get_predictions(logits, d).sum().item() == 2 * 240 * 320
    other = torch.zeros(2, 3, 240, 320)
    with pytest.raises(RuntimeError):
        egn_b200.get_predictions(other, d)          # different storage -> argmax kernel path (CUDA only)
    logits.add_(1)                                   # same storage, new version
    with pytest.raises(RuntimeError):
        egn_b200.get_predictions(logits, d)
