#!/bin/bash
# Profiling pass of one round (run under gpurun on one B200).  Raw ncu reports are converted to CSV
# pages on the box and deleted (gpurun_out/ is capped at 64 MiB); tools/summarize_profiles.py turns
# the CSVs into the tracked summaries under profiles/.
#   gpurun --timeout 900 -- 'bash tools/profile_round.sh r01'
R=${1:-r01}
O=gpurun_out
mkdir -p $O
# the 16-frame captures must run the tiling bench.py uses at its 128-frame micro-batch, not the latency-mode one
export EGN_TC_NO_LATENCY_MODE=1
full() {   # name, kernel regex, skip, count, command...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $skip -c $cnt -o $O/${R}_$name -f "$@" > /dev/null 2>&1
  ncu -i $O/${R}_$name.ncu-rep --page raw --csv > $O/${R}_${name}_raw.csv 2>/dev/null
  ncu -i $O/${R}_$name.ncu-rep --page source --csv --launch-skip 0 --launch-count 1 > $O/${R}_${name}_source0.csv 2>/dev/null
  rm -f $O/${R}_$name.ncu-rep
}
# 1. every launch of bench.py's own command (warm-up step + the two timed steps = 3 x 2 micro-batches of 128 frames)
#    with its device time (cold-cache, serialised: compare shares with roofline.kernel_share_of_step, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 960 --csv --log-file $O/${R}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/${R}_launches.log 2>&1
# 2. DRAM traffic + duration of every conv_tc launch of one warm pass per net
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:conv_tc -s 38 -c 38 \
    --csv --log-file $O/${R}_conv_dram_bdcn.csv python tools/profile_layer.py bdcn 16 > /dev/null 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:conv_tc -s 54 -c 54 \
    --csv --log-file $O/${R}_conv_dram_esf.csv python tools/profile_layer.py esf 16 > /dev/null 2>&1
# 3. full captures: first six BDCN conv launches of the warm pass (msblock1_1.conv, .tail, conv1_2, msblock1_2.conv, .tail,
#    conv2_1), conv3_2 (launch 14), and the first seven ESF conv launches (54 conv_tc launches per ESF pass: dec.final.conv2 and the four half-resolution decoder pre-convolutions included) (head.conv2, down_block1.{conv1,conv21,conv22,conv31,conv32,TD})
full conv_bdcn conv_tc 38 6 python tools/profile_layer.py bdcn 16
full conv_vgg3_2 conv_tc 52 1 python tools/profile_layer.py bdcn 16
full conv_esf conv_tc 54 7 python tools/profile_layer.py esf 16
# 4. the bandwidth-bound helpers of one warm pass
full aux_esf "instnorm|first_conv|head_tail|spatial_mean" 14 14 python tools/profile_layer.py esf 16
full aux_bdcn "first_conv|maxpool|bdcn_tail" 6 6 python tools/profile_layer.py bdcn 16
du -sh $O; ls $O | head -30
