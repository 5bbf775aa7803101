#!/usr/bin/env python
"""Headline benchmark: frames/s of the edge (BDCN) + ESF-Net forward at 240x320.

    python bench.py --gpus N --steps K --warmup W            # the B200 engine (libegn.so)
    python bench.py --impl reference --steps K --warmup W    # CPU arm: the oracle port on host cores

Workload (BASELINE.json configs[1]): configs/baseline_edge.yaml, batch 256 per GPU, synthetic
240x320 frames (seeded z-scored noise), synthetic checkpoints (oracle/synth.py).
One step = calc_edge + DenseNet2D forward + argmax / soft-argmax centres + metric accumulation for
one batch.  Prints ONE JSON line (rank 0).

    python bench.py --config baseline_adain_edge ...         # BASELINE.json configs[2]
    python bench.py --stream 8 --steps 200                   # configs[3]: evaluate.py's streaming path at batch 8
                                                             # (u8 ingest -> edge -> ESF-Net -> argmax -> ellipse fit
                                                             #  -> D2H), latency p50 / p99 in the same JSON schema

The batch carries PLANTED frames (tests/golden/planted_<config>.npz: the reference's own outputs for two
golden frames and eight real eye crops, at batch positions 0 / 127 / 255 and spread through the batch);
after the timed region their argmax agreement, centre error and ellipse-parameter error go into the JSON
line ("parity"), so the benchmarked configuration itself (micro-batch, tiling) is pinned to the reference.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GFLOP_PER_FRAME = {"baseline_edge": 129.603, "baseline": 25.455 + 83.524, "baseline_adain_edge": 144.261}
METRIC = "frames/sec @240x320 edge+ESF-Net fwd"


def read_peaks():
    """(bf16 TFLOP/s, HBM GB/s, source): the driver-written MEASURED_PEAKS.json (sustained figures: the
    kernels are timed inside a long step), else the fallback of B200_PROFILING.md."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    tf, gbs, src = 1400.0, 6650.0, "fallback (B200_PROFILING.md)"
    if os.path.isfile(p):
        try:
            d = json.load(open(p))

            def pick(keys):
                for k in keys:
                    v = d.get(k)
                    if isinstance(v, dict):
                        v = v.get("sustained", v.get("value"))
                    if isinstance(v, (int, float)) and v > 0:
                        return float(v)
                return None
            t = pick(["bf16_tflops_sustained", "bf16_tflops", "bf16_tflops_burst", "tensor_tflops"])
            g = pick(["hbm_gbs", "hbm_gbs_sustained", "hbm_gb_s", "hbm_gbps"])
            if t:
                tf, src = t, "measured (MEASURED_PEAKS.json)"
            if g:
                gbs = g
        except Exception:
            pass
    return tf, gbs, src


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for i, n in enumerate(names):
                if len(r) > 2 + i and r[2 + i].lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}



PLANTED_CONFIGS = ("baseline_edge", "baseline_adain_edge")


def planted_frames():
    """The ten frames of oracle/make_golden_planted.py (same construction; fixtures only, no oracle code)."""
    import numpy as np
    gd = os.path.join(ROOT, "tests", "golden")
    img = np.load(os.path.join(gd, "fwd_input.npz"))["img"]
    fr = np.load(os.path.join(gd, "frames_u8.npz"))["frames"][:8].astype(np.float64)
    z = np.stack([((f - f.mean()) / f.std()).astype(np.float32)[None] for f in fr])
    return np.concatenate([img, z], 0)


def planted_positions(B):
    """[(batch position, planted index)]: golden frame 0 at 0 and B-1, golden frame 1 at B/2-1, the eight real
    crops spread through the batch."""
    if B < 16:
        return []
    pos = [(0, 0), (B // 2 - 1, 1), (B - 1, 0)]
    pos += [(int((k + 0.5) * B / 8), 2 + k) for k in range(8)]
    return pos


def plant(x_host, B):
    import torch
    pos = planted_positions(B)
    if pos:
        pf = torch.from_numpy(planted_frames())
        for p, i in pos:
            x_host[p] = pf[i]
    return pos


def planted_parity(config, pos, argmax, el_pred, el_out):
    """Planted frames of the batch against the reference's outputs (north_star bars: >= 99.9 % argmax
    agreement, centres within 0.25 px, ellipse parameters within 1e-2 relative with floor 1e-2)."""
    import numpy as np
    path = os.path.join(ROOT, "tests", "golden", "planted_%s.npz" % config)
    if not pos or not os.path.isfile(path):
        return None
    g = np.load(path)
    am = argmax.cpu().numpy()
    ep, eo = el_pred.cpu().numpy().astype(np.float64), el_out.cpu().numpy().astype(np.float64)
    scale = np.array([160.0, 120.0])
    agree, centre, ell = 1.0, 0.0, 0.0
    par = [2, 3, 4, 7, 8, 9]
    for p, i in pos:
        agree = min(agree, float((am[p] == g["pred"][i]).mean()))
        for sl in (slice(0, 2), slice(5, 7)):
            centre = max(centre, float(np.abs((ep[p, sl] - g["elPred"][i, sl]) * scale).max()),
                         float(np.abs((eo[p, sl] - g["elOut"][i, sl]) * scale).max()))
        for got, want in ((ep[p, par], g["elPred"][i, par]), (eo[p, par], g["elOut"][i, par])):
            ell = max(ell, float(np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-2))))
    ok = agree >= 0.999 and centre < 0.25 and ell < 1e-2
    return {"planted_frames": len(pos), "positions": [p for p, _ in pos], "argmax_agreement": agree,
            "centre_px": centre, "ell_rel": ell, "pass": bool(ok),
            "against": "tests/golden/planted_%s.npz (outputs of the unmodified reference, oracle/make_golden_planted.py)" % config}


def dtype_label(info):
    if info.get("lowered_layers"):
        return "bf16x3 with %d layers lowered to two products (EGN_PRODUCTS probe: NOT the parity engine)" % info["lowered_layers"]
    if not info["tensor_core_path"]:
        return "f32-simt (EGN_CONV=simt debugging path: NOT the benchmarked engine)"
    if info["products_per_mac"] == 3:
        return "bf16x3 (split-bf16 operands: hi*hi + lo*hi + hi*lo tcgen05 products per MAC, fp32 accumulate)"
    return "bf16x%d (EGN_NSPLIT=%d: does NOT meet the parity bars)" % (info["products_per_mac"], info["products_per_mac"])


def cpu_reference_fps(frames, repeats, threads, config="baseline_edge"):
    """The CPU arm: oracle/graph.py (torch fp32 port of the reference path) on host cores."""
    import torch
    from oracle import graph, synth
    torch.set_num_threads(threads)
    st = synth.SETTINGS[config]
    bsd, esd = synth.make_bdcn_state(0), synth.make_esf_state(st, 0)
    img = synth.randn_frames(frames, seed=1)
    best = None
    with torch.no_grad():
        for _ in range(repeats):
            t0 = time.time()
            edge = graph.calc_edge(bsd, img)
            out = graph.esf_forward(esd, st, img, edge)
            graph.get_predictions(out["op"])
            dt = time.time() - t0
            best = dt if best is None else min(best, dt)
    return frames / best, best


def workload_config(config, batch, world, micro_batch):
    return {"workload": "%s.yaml: BDCN edge extractor + ESF-Net, batch %d per GPU, 240x320" % (config, batch),
            "global_batch": batch * world, "micro_batch": micro_batch,
            "parallelism": "dp%d (batch-sharded frames)" % world,
            "cache": "activations written and re-read by each step (tens of GB at this micro-batch) exceed the 126 MB L2; no flush needed",
            "weights": "synthetic seeded checkpoints in the reference container formats"}


def conv_traffic(config, micro_batch):
    """Per-launch DRAM traffic of conv_tc_kernel from the newest committed ncu capture (profiles/): a capture
    taken at this micro-batch and configuration is used as is, anything else is scaled and SAYS so."""
    import glob
    best = None
    for p in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_conv_traffic*.json"))):
        try:
            d = json.load(open(p))
        except Exception:
            continue
        d["file"] = os.path.basename(p)
        exact = int(d.get("frames", 16)) == micro_batch and d.get("config", "baseline_edge") == config
        if best is None or exact or not best[0]:
            best = (exact, d)
    return best


def run_reference(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    frames = args.cpu_frames
    from oracle import graph, synth
    torch.set_num_threads(threads)
    st = synth.SETTINGS[args.config]
    bsd, esd = synth.make_bdcn_state(0), synth.make_esf_state(st, 0)
    img = synth.randn_frames(frames, seed=1)

    def step():
        with torch.no_grad():
            edge = graph.calc_edge(bsd, img)
            out = graph.esf_forward(esd, st, img, edge)
            graph.get_predictions(out["op"])
    for _ in range(args.warmup):
        step()
    t0 = time.time()
    for _ in range(args.steps):
        step()
    dt = time.time() - t0
    fps = frames * args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(args.config, args.batch, max(1, args.gpus), args.micro_batch),
                           sample="%d frames per step on the host CPU (bounded sample of the %d-frame batch)" % (frames, args.batch)),
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port",
                             "sample": "%d steps x %d frames, oracle/graph.py (torch fp32), all host threads" % (args.steps, frames)},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def run_stream(args):
    """BASELINE.json configs[3]: evaluate.py's per-frame video path (evaluate.py:235-285) at `--stream` frames
    per call: uint8 frames on the host -> H2D -> z-score ingest -> BDCN edge -> ESF-Net -> argmax / centres ->
    both ellipses refined -> D2H of edge map, segmentation and ellipses.  One step = one such call; `value`
    is frames/s with the call's inputs resident on the device, `e2e` the same from pinned host memory to
    host results; latency percentiles of the e2e call are reported beside them."""
    import numpy as np
    import torch
    import egn_b200
    from oracle import synth           # synthetic checkpoints only (test infrastructure)
    assert args.gpus == 1, "the streaming configuration is a single-GPU latency measurement"
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    B = args.stream
    st = synth.SETTINGS[args.config]
    edge_model = egn_b200.BDCN(); edge_model.load_state_dict(synth.make_bdcn_state(0))
    model = egn_b200.DenseNet2D(st); model.load_state_dict(synth.make_esf_state(st, 0))
    edge_model = edge_model.to(dev).eval(); model = model.to(dev).eval()
    edge_model.micro_batch = model.micro_batch = B
    fr = np.load(os.path.join(ROOT, "tests", "golden", "frames_u8.npz"))["frames"]
    frames = torch.from_numpy(np.stack([fr[i % len(fr)] for i in range(B)])).pin_memory()
    frames_dev = frames.to(dev)
    h_edge = torch.empty((B, 1, 240, 320), dtype=torch.float32).pin_memory()
    h_seg = torch.empty((B, 240, 320), dtype=torch.uint8).pin_memory()
    h_ell = torch.empty((B, 2, 5), dtype=torch.float64).pin_memory()
    ctx = None

    def call(src):
        x = egn_b200.preprocess_frames_u8(src, dev)
        edge = edge_model.edge(x)
        logits, el_out, latent, argmax, el_pred = model.infer(x, edge, None)
        ell = model.context(dev).ellipse_refine(argmax, el_pred, True)
        return edge, argmax, ell

    def step_resident():
        call(frames_dev)

    def step_e2e():
        edge, argmax, ell = call(frames.to(dev, non_blocking=True))
        h_edge.copy_(edge, non_blocking=True); h_seg.copy_(argmax, non_blocking=True); h_ell.copy_(ell, non_blocking=True)
        torch.cuda.synchronize(dev)

    with torch.no_grad():
        for _ in range(max(3, args.warmup)):
            step_e2e()
        ectx, mctx = edge_model.context(dev), model.context(dev)
        l0 = ectx.launch_count() + mctx.launch_count()
        sampler = ClockSampler(0); sampler.start()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step_resident()
        e1.record(); torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1)
        launches = ectx.launch_count() + mctx.launch_count() - l0
        lat = []
        for _ in range(args.steps):
            t0 = time.perf_counter()
            step_e2e()
            lat.append((time.perf_counter() - t0) * 1e3)
        sampler.stop_flag = True
        info = mctx.info()
    lat.sort()
    pct = lambda q: lat[min(len(lat) - 1, int(q * len(lat)))]
    ms_e2e = sum(lat)
    gf = GFLOP_PER_FRAME.get(args.config)
    peak_tf, peak_gbs, peak_src = read_peaks()
    fps = B * args.steps / (ms / 1e3)
    line = {"metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": 1, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": dtype_label(info), "data": "real eye crops (tests/golden/frames_u8.npz, videos/example1.avi) tiled to the batch; synthetic weights",
            "config": {"workload": "evaluate.py video path (%s.yaml): streaming 240x320 frames at batch %d, latency plus ellipse fit, 1 B200"
                                   % (args.config, B), "global_batch": B, "micro_batch": B, "parallelism": "dp1",
                       "cache": "latency measurement: one small call per step, L2-resident by nature (that is the workload)"},
            "latency_ms": {"p50": pct(0.5), "p99": pct(0.99), "mean": ms_e2e / len(lat), "calls": len(lat),
                           "what": "host u8 frames -> edge map, segmentation and both refined ellipses on the host, per call"},
            "e2e": {"value": B * len(lat) / (ms_e2e / 1e3), "unit": "frames/s", "h2d_bytes_per_step": int(frames.numel()),
                    "d2h_bytes_per_step": int(h_edge.numel() * 4 + h_seg.numel() + h_ell.numel() * 8), "ms_per_step": ms_e2e / len(lat)},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "kernel": "whole call (latency-bound at this batch)", "achieved": fps * gf / 1000.0 if gf else None,
                         "peak": peak_tf, "unit": "TFLOP/s", "frac": fps * gf / 1000.0 / peak_tf if gf else None, "peak_source": peak_src,
                         "traffic": None},
            "clocks": sampler.summary()}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="egn", choices=["egn", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="frames per GPU per step")
    ap.add_argument("--micro-batch", type=int, default=int(os.environ.get("EGN_MICRO_BATCH", "256")))
    ap.add_argument("--config", default="baseline_edge")
    ap.add_argument("--cpu-frames", type=int, default=32, help="frames per CPU step (bounded sample of the workload: ~8 s per step on 16 cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--layer-table", default=None, help="write the per-layer kernel-time CSV of the timed region here")
    ap.add_argument("--stream", type=int, default=0, help="streaming mode (evaluate.py video path): frames per call, e.g. 1 / 8 / 32")
    ap.add_argument("--allow-nonparity", action="store_true", help="accept EGN_NSPLIT / EGN_CONV knobs that fail the parity bars (the line is labelled)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.stream:
        return run_stream(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    import egn_b200
    from oracle import synth           # synthetic checkpoints / inputs only (test infrastructure)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL may print a version banner to stdout when its communicator is created; stdout must carry
        # exactly one JSON line, so fd 1 points at stderr until the communicator exists
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            warm = torch.zeros(1, device=dev)
            dist.all_reduce(warm)
            torch.cuda.synchronize(dev)
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    assert args.gpus == world, "--gpus must equal WORLD_SIZE (launch N>1 with torch.distributed.run)"

    st = synth.SETTINGS[args.config]
    # the step drives both modules in stream order on one stream, so their activation arenas may be one pool
    egn_b200.share_workspace(not os.environ.get("EGN_NO_SHARE_WORKSPACE"))
    edge_model = egn_b200.BDCN(); edge_model.load_state_dict(synth.make_bdcn_state(0))
    model = egn_b200.DenseNet2D(st); model.load_state_dict(synth.make_esf_state(st, 0))
    edge_model = edge_model.to(dev).eval(); model = model.to(dev).eval()
    edge_model.micro_batch = model.micro_batch = args.micro_batch
    B = args.batch
    # weak scaling: every rank owns its own B frames (frames shard by batch, no data-path collective)
    x_host = synth.randn_frames(B, seed=100 + rank)
    planted = plant(x_host, B) if args.config in PLANTED_CONFIGS else []
    x_host = x_host.pin_memory()
    lab_host = synth.evaluate_style_labels(B).to(torch.uint8).pin_memory()
    x_dev = x_host.to(dev)
    lab_dev = lab_host.to(dev)
    cond = torch.zeros(B, 4, device=dev)
    centres = torch.full((B, 2), 100.0, device=dev)
    acc = egn_b200.MetricAccumulator(dev)

    def step_resident():
        edge = edge_model.edge(x_dev)
        logits, el_out, latent, argmax, el_pred = model.infer(x_dev, edge, cond)
        model.context(dev).metrics_accumulate(argmax, lab_dev, cond, acc.acc, centres, centres, el_out, el_pred)
        return argmax, el_pred

    out_am = torch.empty((B, 240, 320), dtype=torch.uint8).pin_memory()
    out_el = torch.empty((B, 10), dtype=torch.float32).pin_memory()
    out_eo = torch.empty((B, 10), dtype=torch.float32).pin_memory()
    out_lat = torch.empty((B, 153), dtype=torch.float32).pin_memory()

    def step_e2e():
        # the call a user of the reference makes: CPU tensors in (test.py:78-92), CPU results out
        xd = x_host.to(dev, non_blocking=True)
        edge = egn_b200.calc_edge(None, xd, edge_model, dev)
        op, el_pred, latent, _, el_out = model(xd, edge, None, None, None, None, None, cond, 0, 0)
        out_am.copy_(model.last_argmax, non_blocking=True)
        out_el.copy_(el_pred, non_blocking=True); out_eo.copy_(el_out, non_blocking=True)
        out_lat.copy_(latent, non_blocking=True)
        torch.cuda.synchronize(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        mine = float(e0.elapsed_time(e1))
        per_rank = [mine]
        if world > 1:
            allms = torch.zeros(world, device=dev, dtype=torch.float64)
            allms[rank] = mine
            dist.all_reduce(allms, op=dist.ReduceOp.SUM)
            per_rank = [float(v) for v in allms.tolist()]
        timed.per_rank = per_rank
        return max(per_rank)

    with torch.no_grad():
        for _ in range(args.warmup):
            step_resident()
        ectx, mctx = edge_model.context(dev), model.context(dev)
        l0 = ectx.launch_count() + mctx.launch_count()
        ectx.profile(True); mctx.profile(True)
        ectx.profile_read(True); mctx.profile_read(True)
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        ms = timed(step_resident, args.steps)
        rank_ms = [v / args.steps for v in timed.per_rank]
        acc.all_reduce()
        sampler.stop_flag = True
        launches = ectx.launch_count() + mctx.launch_count() - l0
        pe, pm = ectx.profile_read(True), mctx.profile_read(True)
        if args.layer_table and rank == 0:
            with open(args.layer_table, "w") as f:
                f.write(ectx.profile_table())
                f.write(mctx.profile_table())
        ectx.profile(False); mctx.profile(False)
        for _ in range(2):
            step_e2e()
        ms_e2e = timed(step_e2e, args.steps)
        # the planted frames of the LAST end-to-end step (host results of the public call path)
        parity = planted_parity(args.config, planted, out_am, out_el, out_eo)
        info = mctx.info()
        einfo = ectx.info()
    if (info["products_per_mac"] != 3 or not info["tensor_core_path"] or einfo["products_per_mac"] != 3
            or not einfo["tensor_core_path"] or info["lowered_layers"] or einfo["lowered_layers"]) and not args.allow_nonparity:
        raise SystemExit("bench.py: the engine runs with EGN_NSPLIT=%d / tensor_core_path=%d, which is not the parity "
                         "configuration; pass --allow-nonparity to time it anyway (the line is labelled)"
                         % (info["products_per_mac"], info["tensor_core_path"]))

    fps = world * B * args.steps / (ms / 1000.0)
    fps_e2e = world * B * args.steps / (ms_e2e / 1000.0)
    conv_ms, conv_flops, conv_n = pe[0] + pm[0], pe[1] + pm[1], pe[2] + pm[2]
    peak_tf, peak_gbs, peak_src = read_peaks()
    achieved = conv_flops / (conv_ms / 1000.0) / 1e12 if conv_ms > 0 else 0.0
    gf = GFLOP_PER_FRAME.get(args.config)
    tr = conv_traffic(args.config, min(args.micro_batch, B))
    traffic = tr[1] if tr else None
    products = info["products_per_mac"]
    if rank == 0:
        line = {"metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": dtype_label(dict(info, lowered_layers=info["lowered_layers"] + einfo["lowered_layers"])), "data": "synthetic",
                "config": workload_config(args.config, B, world, args.micro_batch),
                "e2e": {"value": fps_e2e, "unit": "frames/s", "h2d_bytes_per_step": int(x_host.numel() * 4),
                        "d2h_bytes_per_step": int(out_am.numel() + 4 * (out_el.numel() + out_eo.numel() + out_lat.numel())),
                        "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": int(launches),
                "roofline": {"bound": "tensor", "kernel": "conv_tc_kernel (all %d conv launches of the timed region)" % conv_n,
                             "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                             "peak_source": peak_src,
                             # every algorithmic MAC costs `products` bf16 tensor-core products, so
                             # frac can never exceed 1/products; the second figure is the fraction of THAT ceiling
                             "products_per_mac": products, "policy_ceiling_frac": 1.0 / products,
                             "frac_of_policy_ceiling": achieved / peak_tf * products,
                             "traffic": ((traffic["dram_bytes_per_launch"] * min(args.micro_batch, B) / float(traffic.get("frames", 16)))
                                         if traffic else None),
                             "traffic_note": ((traffic["note"] + (" [%s]" % traffic["file"]) +
                                               ("" if tr[0] else "; EXTRAPOLATED: scaled from that capture to %d-frame launches of %s"
                                                % (min(args.micro_batch, B), args.config)))
                                              if traffic else "no ncu capture committed yet"),
                             "algorithmic_flops_per_launch": conv_flops / conv_n if conv_n else None,
                             "ms_per_launch": conv_ms / conv_n if conv_n else None,
                             "kernel_share_of_step": conv_ms / ms if ms > 0 else None,
                             "whole_step_tflops": fps / world * gf / 1000.0 if gf else None,
                             "whole_step_frac": fps / world * gf / 1000.0 / peak_tf if gf else None},
                "clocks": sampler.summary(),
                "hbm_used_gb": round((torch.cuda.mem_get_info(dev)[1] - torch.cuda.mem_get_info(dev)[0]) / 1e9, 1),
                "metrics_check": acc.result()["frames"], "parity": parity,
                "rank_ms_per_step": rank_ms,
                "workspace_gb": round((info["workspace_bytes"] + einfo["workspace_bytes"] + max(info["shared_pool_bytes"], einfo["shared_pool_bytes"])) / 1e9, 2),
                "workspace_note": "one activation pool per device shared by the BDCN and ESF-Net contexts (egn_share_workspace: the evaluator runs them in stream order)"}
        if gf:
            line["roofline"]["whole_step_frac_of_policy_ceiling"] = fps / world * gf / 1000.0 / peak_tf * products
        if not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            cfps, csec = cpu_reference_fps(args.cpu_frames, 2, threads, args.config)
            line["cpu_baseline"] = {"value": cfps, "unit": "frames/s", "cores": threads, "kind": "port",
                                    "sample": "best of 2 x %d frames (%.1f s each), oracle/graph.py torch-fp32 port of "
                                              "calc_edge + DenseNet2D forward + get_predictions" % (args.cpu_frames, csec)}
        line["metrics_check"] = int(line["metrics_check"])
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
