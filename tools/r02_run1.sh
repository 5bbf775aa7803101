#!/bin/bash
# r02 GPU pass 1: GPU tests, baseline bench of the round-1 kernels with the new bench line, sanitizer records
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -x -q -s > $O/r02a_tests.log 2>&1; echo "tests rc=$?" >> $O/r02a_tests.log
tail -5 $O/r02a_tests.log
python bench.py --steps 10 --warmup 3 --layer-table $O/r02a_layers.csv > $O/r02a_bench.json 2> $O/r02a_bench.err; tail -c 1500 $O/r02a_bench.json
for tool in racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 5 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02a_$tool.log 2>&1
  tail -4 $O/r02a_$tool.log
done
