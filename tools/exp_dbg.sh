mkdir -p gpurun_out
EGN_TC_TIMING=1 python tools/profile_layer.py esf 16 2> gpurun_out/x1_timing_esf.log >/dev/null
EGN_TC_TIMING=1 python tools/profile_layer.py bdcn 16 2> gpurun_out/x1_timing_bdcn.log >/dev/null
for d in 0 1 2 4 8 6 12; do
  EGN_TC_DBG=$d python bench.py --batch 64 --micro-batch 64 --steps 2 --warmup 3 --no-cpu-baseline --layer-table gpurun_out/x1_layers_dbg$d.csv > gpurun_out/x1_bench_dbg$d.json 2>/dev/null
done
ls gpurun_out | head -40
