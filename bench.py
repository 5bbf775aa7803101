#!/usr/bin/env python
"""Headline benchmark: frames/s of the edge (BDCN) + ESF-Net forward at 240x320.

    python bench.py --gpus N --steps K --warmup W            # the B200 engine (libegn.so)
    python bench.py --impl reference --steps K --warmup W    # CPU arm: the oracle port on host cores

Workload (BASELINE.json configs[1]): configs/baseline_edge.yaml, batch 256 per GPU, synthetic
240x320 frames (seeded z-scored noise), synthetic checkpoints (oracle/synth.py).
One step = calc_edge + DenseNet2D forward + argmax / soft-argmax centres + metric accumulation for
one batch.  Prints ONE JSON line (rank 0).
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


def cpu_reference_fps(frames, repeats, threads):
    """The CPU arm: oracle/graph.py (torch fp32 port of the reference path) on host cores."""
    import torch
    from oracle import graph, synth
    torch.set_num_threads(threads)
    st = synth.SETTINGS["baseline_edge"]
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
            "cache": "working set per step (%.1f GB activations) exceeds the 126 MB L2; no flush needed" % (micro_batch * 0.4),
            "weights": "synthetic seeded checkpoints in the reference container formats"}


def conv_traffic():
    """Per-launch DRAM traffic of conv_tc_kernel from the committed ncu capture (profiles/)."""
    p = os.path.join(ROOT, "profiles", "r01_conv_traffic.json")
    if os.path.isfile(p):
        try:
            return json.load(open(p))
        except Exception:
            return None
    return None


def run_reference(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    frames = args.cpu_frames
    from oracle import graph, synth
    torch.set_num_threads(threads)
    st = synth.SETTINGS["baseline_edge"]
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
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

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
    edge_model = egn_b200.BDCN(); edge_model.load_state_dict(synth.make_bdcn_state(0))
    model = egn_b200.DenseNet2D(st); model.load_state_dict(synth.make_esf_state(st, 0))
    edge_model = edge_model.to(dev).eval(); model = model.to(dev).eval()
    edge_model.micro_batch = model.micro_batch = args.micro_batch
    B = args.batch
    # weak scaling: every rank owns its own B frames (frames shard by batch, no data-path collective)
    x_host = synth.randn_frames(B, seed=100 + rank).pin_memory()
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
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

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

    fps = world * B * args.steps / (ms / 1000.0)
    fps_e2e = world * B * args.steps / (ms_e2e / 1000.0)
    conv_ms, conv_flops, conv_n = pe[0] + pm[0], pe[1] + pm[1], pe[2] + pm[2]
    peak_tf, peak_gbs, peak_src = read_peaks()
    achieved = conv_flops / (conv_ms / 1000.0) / 1e12 if conv_ms > 0 else 0.0
    gf = GFLOP_PER_FRAME.get(args.config)
    traffic = conv_traffic()
    if rank == 0:
        line = {"metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16x3 (split-bf16 operands, fp32 accumulate)", "data": "synthetic",
                "config": workload_config(args.config, B, world, args.micro_batch),
                "e2e": {"value": fps_e2e, "unit": "frames/s", "h2d_bytes_per_step": int(x_host.numel() * 4),
                        "d2h_bytes_per_step": int(out_am.numel() + 4 * (out_el.numel() + out_eo.numel() + out_lat.numel())),
                        "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": int(launches),
                "roofline": {"bound": "tensor", "kernel": "conv_tc_kernel (all %d conv launches of the timed region)" % conv_n,
                             "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                             "peak_source": peak_src,
                             # the capture ran 16-frame launches; a launch of this run carries micro_batch frames
                             "traffic": ((traffic["dram_bytes_per_launch"] * min(args.micro_batch, B) / float(traffic.get("frames", 16)))
                                         if traffic else None),
                             "traffic_note": ((traffic["note"] + "; scaled here to %d-frame launches" % min(args.micro_batch, B))
                                              if traffic else "no ncu capture committed yet"),
                             "algorithmic_flops_per_launch": conv_flops / conv_n if conv_n else None,
                             "ms_per_launch": conv_ms / conv_n if conv_n else None,
                             "kernel_share_of_step": conv_ms / ms if ms > 0 else None,
                             "whole_step_tflops": fps / world * gf / 1000.0 if gf else None,
                             "whole_step_frac": fps / world * gf / 1000.0 / peak_tf if gf else None},
                "clocks": sampler.summary(),
                "hbm_used_gb": round((torch.cuda.mem_get_info(dev)[1] - torch.cuda.mem_get_info(dev)[0]) / 1e9, 1),
                "metrics_check": acc.result()["frames"]}
        if not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            cfps, csec = cpu_reference_fps(args.cpu_frames, 2, threads)
            line["cpu_baseline"] = {"value": cfps, "unit": "frames/s", "cores": threads, "kind": "port",
                                    "sample": "best of 2 x %d frames (%.1f s each), oracle/graph.py torch-fp32 port of "
                                              "calc_edge + DenseNet2D forward + get_predictions" % (args.cpu_frames, csec)}
        line["metrics_check"] = int(line["metrics_check"])
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
