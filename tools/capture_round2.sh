#!/bin/bash
# Round-2 capture, run on the GPU box:  bash tools/capture_round2.sh r2h
# Leaves everything under gpurun_out/<tag>_*; numbers printed under ncu are never bench values.
tag=${1:-r2x}
out=gpurun_out
mkdir -p $out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $out/${tag}_launches_bf16x3.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-inference > $out/${tag}_launches.log 2>&1
python tools/launches.py $out/${tag}_launches_bf16x3.csv $out/${tag}_launches_bf16x3_summary.csv > $out/${tag}_launches_digest.txt 2>&1
# one full-set capture of the training step's tensor-core kernels (coarse + fine launches of one step)
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"field_dgrad_pipe_kernel|field_wgrad_kernel|field_fwd_pipe_kernel" \
  --launch-skip 18 -c 6 -o $out/${tag}_train_bf16x3 -f python tools/step_breakdown.py bf16x3 > $out/${tag}_ncu_train.log 2>&1
python tools/ncu_top.py $out/${tag}_train_bf16x3.ncu-rep 20 > $out/${tag}_train_bf16x3_ncu_summary.txt 2>&1
python tools/ncu_traffic.py $out/${tag}_train_bf16x3.ncu-rep bf16x3/train > $out/${tag}_traffic.log 2>&1
# NeRF++: the foreground and background networks' kernels (pipelined forward XS=4/6, pipelined dgrad XN=64/96) and the shared wgrad
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"field_fwd_pipe_kernel|field_wgrad_kernel|field_dgrad_pipe_kernel" \
  --launch-skip 24 -c 12 -o $out/${tag}_pp_bf16x3 -f python tools/pp_step.py bf16x3 > $out/${tag}_ncu_pp.log 2>&1
python tools/ncu_top.py $out/${tag}_pp_bf16x3.ncu-rep 12 > $out/${tag}_pp_bf16x3_ncu_summary.txt 2>&1
python tools/ncu_traffic.py $out/${tag}_pp_bf16x3.ncu-rep bf16x3/nerfpp-train >> $out/${tag}_traffic.log 2>&1
cp profiles/traffic.json $out/${tag}_traffic.json
ls -la $out/*.ncu-rep; rm -f $out/${tag}_train_bf16x3.ncu-rep $out/${tag}_pp_bf16x3.ncu-rep
cat $out/${tag}_launches_digest.txt; cat $out/${tag}_traffic.log; tail -n 3 $out/${tag}_ncu_train.log $out/${tag}_ncu_pp.log
