mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --batch 256 --micro-batch 128 --steps 3 --warmup 3 --no-cpu-baseline --layer-table gpurun_out/x10_layers.csv > gpurun_out/x10_bench.json 2>/dev/null
EGN_LAST_SIMT=1 python bench.py --batch 256 --micro-batch 128 --steps 3 --warmup 3 --no-cpu-baseline --layer-table gpurun_out/x10_layers_lastsimt.csv > gpurun_out/x10_bench_lastsimt.json 2>/dev/null
python bench.py --batch 256 --micro-batch 256 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/x10_bench_mb256.json 2>gpurun_out/x10_mb256.err
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
for f in x10_bench x10_bench_lastsimt x10_bench_mb256; do cut -c1-110 gpurun_out/$f.json; done; grep -h "final.conv2\|last_conv" gpurun_out/x10_layers*.csv; tail -2 gpurun_out/x10_mb256.err
