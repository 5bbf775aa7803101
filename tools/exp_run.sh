mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tools/stream_bench.py --out gpurun_out/x4_stream.json > gpurun_out/x4_stream.log 2>&1
python tools/stream_bench.py --no-refine --out gpurun_out/x4_stream_norefine.json > gpurun_out/x4_stream_norefine.log 2>&1
python bench.py --batch 64 --micro-batch 64 --steps 2 --warmup 3 --no-cpu-baseline --layer-table gpurun_out/x4_layers.csv > gpurun_out/x4_bench.json 2>/dev/null
cut -c1-200 gpurun_out/x4_stream.log; cut -c1-200 gpurun_out/x4_stream_norefine.log; cut -c1-120 gpurun_out/x4_bench.json
