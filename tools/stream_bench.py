"""Streaming (video) latency of the per-image path, BASELINE.json configs[3] / SURVEY.md 8d C4:

    python tools/stream_bench.py [--batches 1,8,32] [--iters 60] [--config baseline_edge] [--out file.json]

Each iteration is what evaluate.py:244-251 does per eye crop, batched B crops at a time: uint8
240x320 frames in pinned host memory -> H2D -> per-frame z-score (evaluate.py:102-103) -> BDCN edge
-> ESF-Net forward -> argmax -> normalised->pixel ellipses + the IoU refinement of both ellipses
(evaluate.py:141-151) -> D2H of the edge map, the segmentation map and the two ellipses.  Latency is
host wall clock around the whole call (synchronised both sides); p50 / p99 over --iters calls after
warm-up.  One JSON line per batch size on stdout; the same list is written to --out."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import egn_b200
from oracle import synth          # synthetic checkpoints / frames only (test infrastructure)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", default="1,8,32")
    ap.add_argument("--iters", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--config", default="baseline_edge")
    ap.add_argument("--no-refine", action="store_true")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    st = synth.SETTINGS[args.config]
    edge_model = egn_b200.BDCN(); edge_model.load_state_dict(synth.make_bdcn_state(0))
    model = egn_b200.DenseNet2D(st); model.load_state_dict(synth.make_esf_state(st, 0))
    edge_model = edge_model.to(dev).eval(); model = model.to(dev).eval()
    batches = [int(b) for b in args.batches.split(",")]
    golden = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "frames_u8.npz")
    real = np.load(golden)["frames"] if os.path.isfile(golden) else None
    results = []
    for B in batches:
        # a streaming deployment plans the engine for its own batch size (latency-mode tiling)
        edge_model.micro_batch = model.micro_batch = B
        if real is not None:
            fr = np.stack([real[i % real.shape[0]] for i in range(B)])
        else:
            fr = (np.random.default_rng(0).random((B, 240, 320)) * 255).astype(np.uint8)
        frames = torch.from_numpy(np.ascontiguousarray(fr)).pin_memory()
        o_edge = torch.empty((B, 240, 320), dtype=torch.float32).pin_memory()
        o_seg = torch.empty((B, 240, 320), dtype=torch.uint8).pin_memory()
        o_ell = torch.empty((B, 2, 5), dtype=torch.float64).pin_memory()
        ctx = model.context(dev)

        def call():
            with torch.no_grad():
                x = egn_b200.preprocess_frames_u8(frames.to(dev, non_blocking=True), dev)
                edge = edge_model.edge(x)
                logits, el_out, latent, argmax, el_pred = model.infer(x, edge, None)
                ell = ctx.ellipse_refine(argmax, el_pred, not args.no_refine)
                o_edge.copy_(edge.view(B, 240, 320), non_blocking=True)
                o_seg.copy_(argmax, non_blocking=True)
                o_ell.copy_(ell, non_blocking=True)
            torch.cuda.synchronize(dev)

        for _ in range(args.warmup):
            call()
        l0 = ctx.launch_count() + edge_model.context(dev).launch_count()
        lat = []
        for _ in range(args.iters):
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            call()
            lat.append((time.perf_counter() - t0) * 1e3)
        launches = ctx.launch_count() + edge_model.context(dev).launch_count() - l0
        lat.sort()
        row = {"workload": "evaluate.py video path, %s, batch %d" % (args.config, B), "batch": B, "iters": args.iters,
               "latency_ms_p50": lat[len(lat) // 2], "latency_ms_p99": lat[min(len(lat) - 1, int(len(lat) * 0.99))],
               "latency_ms_min": lat[0], "frames_per_s": B / (lat[len(lat) // 2] / 1e3),
               "ellipse_refinement": not args.no_refine, "launches_per_call": launches / args.iters,
               "h2d_bytes": int(frames.numel()), "d2h_bytes": int(o_edge.numel() * 4 + o_seg.numel() + o_ell.numel() * 8),
               "frames": "real crops of videos/example1.avi (tests/golden/frames_u8.npz)" if real is not None else "synthetic"}
        print(json.dumps(row)); sys.stdout.flush()
        results.append(row)
    if args.out:
        with open(args.out, "w") as f:
            json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
