#!/bin/bash
# Throughput sweep of baseline_edge over the per-GPU batch (BASELINE.json configs[4]):
#   bash tools/sweep.sh <ngpus> <out.jsonl> [batches...]      (run under gpurun [--gpus N])
# One bench.py JSON line per batch size is appended to <out.jsonl>; for N > 1 every rank holds
# `batch` frames (weak scaling) and the metric accumulators are all-reduced over NCCL.
N=${1:-1}; OUT=${2:-gpurun_out/sweep.jsonl}; shift 2
BATCHES=${@:-64 128 256 512 1024 2048 4096}
mkdir -p "$(dirname "$OUT")"; : > "$OUT"
for b in $BATCHES; do
  if [ "$N" = "1" ]; then
    python bench.py --gpus 1 --batch $b --steps 3 --warmup 3 --no-cpu-baseline >> "$OUT" 2>> "$OUT.err"
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus $N --batch $b --steps 3 --warmup 3 --no-cpu-baseline >> "$OUT" 2>> "$OUT.err"
  fi
done
python - "$OUT" <<'P'
import json, sys
for l in open(sys.argv[1]):
    try: d = json.loads(l)
    except Exception: continue
    print("gpus %d  batch/gpu %5d  %8.1f frames/s  e2e %8.1f  ms/step %8.1f" % (d["n_gpus"], d["config"]["global_batch"] // d["n_gpus"], d["value"], d["e2e"]["value"], d["ms_per_step"]))
P
