mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "calc_acc or throughput or metrics" 2>&1 | tail -3
bash tools/sweep.sh 1 gpurun_out/x6_sweep_n1.jsonl 64 256 1024 4096
tail -3 gpurun_out/x6_sweep_n1.jsonl.err
