#!/bin/bash
# Round-2 profiling pass (run under gpurun on one B200); tools/summarize_profiles.py r02 turns the CSVs into profiles/.
#   gpurun --timeout 1500 -- 'bash tools/profile_round2.sh'
R=r02
O=gpurun_out
mkdir -p $O
export EGN_TC_NO_LATENCY_MODE=1      # the small captures must run the throughput tiling bench.py times
full() {   # name, kernel regex, skip, count, command...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $skip -c $cnt -o $O/${R}_$name -f "$@" > /dev/null 2>&1
  ncu -i $O/${R}_$name.ncu-rep --page raw --csv > $O/${R}_${name}_raw.csv 2>/dev/null
  ncu -i $O/${R}_$name.ncu-rep --page source --csv --launch-skip 0 --launch-count 1 > $O/${R}_${name}_source0.csv 2>/dev/null
  rm -f $O/${R}_$name.ncu-rep
}
# 1. every launch of bench.py's own command with its device time (cold-cache, serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/${R}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/${R}_launches.log 2>&1
# 2. DRAM traffic + duration of every conv_tc launch of one warm pass per net AT THE BENCHMARKED MICRO-BATCH (256 frames):
#    BDCN 34 conv_tc launches per pass (4 merged), ESF 54
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:conv_tc -s 34 -c 34 \
    --csv --log-file $O/${R}_conv_dram_bdcn.csv python tools/profile_layer.py bdcn 256 > /dev/null 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:conv_tc -s 54 -c 54 \
    --csv --log-file $O/${R}_conv_dram_esf.csv python tools/profile_layer.py esf 256 > /dev/null 2>&1
# 3. full captures at 16 frames: the first six BDCN conv launches of the warm pass, the merged conv3_2 (launch 34 + 10),
#    the first seven ESF conv launches and up block 1 (launches 54 + 46 .. 49)
full conv_bdcn conv_tc 34 6 python tools/profile_layer.py bdcn 16
full conv_vgg3_2 conv_tc 44 1 python tools/profile_layer.py bdcn 16
full conv_esf conv_tc 54 7 python tools/profile_layer.py esf 16
full conv_dec conv_tc 100 4 python tools/profile_layer.py esf 16
du -sh $O; ls $O | grep ${R}_ | head -30
