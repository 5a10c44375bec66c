#!/bin/bash
# Round capture, run on the GPU box:  bash tools/capture_round.sh r1i
# Leaves everything under gpurun_out/<tag>_*; numbers printed under ncu are never bench values.
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
if [ -z "$SKIP_TESTS" ]; then
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" >> $out/${tag}_smoke.log
fi
timeout 600 python bench.py > $out/${tag}_bench_bf16x3.json 2> $out/${tag}_bench_bf16x3.err
timeout 300 python bench.py --precision bf16 --no-cpu-baseline > $out/${tag}_bench_bf16.json 2> $out/${tag}_bench_bf16.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_bf16x3.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $out/${tag}_launches.log 2>&1
# one full-set capture per tensor-core kernel: inference forward (the roofline kernel), then the training fine pass
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"field_fused_fwd_kernel|field_fwd_pipe_kernel" --launch-skip 3 -c 1 \
  -o $out/${tag}_fwd_infer_bf16x3 -f python tools/prof_field.py bf16x3 > $out/${tag}_ncu_fwd.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"field_fused_dgrad_kernel|field_dgrad_pipe_kernel|field_wgrad_kernel|field_fused_fwd_kernel|field_fwd_pipe_kernel" \
  --launch-skip 16 -c 6 -o $out/${tag}_train_bf16x3 -f python tools/step_breakdown.py bf16x3 > $out/${tag}_ncu_train.log 2>&1
# digest the captures on the box: the reports themselves exceed what gpurun copies back
python tools/ncu_top.py $out/${tag}_fwd_infer_bf16x3.ncu-rep 20 > $out/${tag}_fwd_infer_bf16x3_ncu_summary.txt 2>&1
python tools/ncu_top.py $out/${tag}_train_bf16x3.ncu-rep 20 > $out/${tag}_train_bf16x3_ncu_summary.txt 2>&1
python tools/ncu_traffic.py $out/${tag}_fwd_infer_bf16x3.ncu-rep bf16x3/inference > $out/${tag}_traffic.log 2>&1
python tools/ncu_traffic.py $out/${tag}_train_bf16x3.ncu-rep bf16x3/train >> $out/${tag}_traffic.log 2>&1
cp profiles/traffic.json $out/${tag}_traffic.json
ncu -i $out/${tag}_train_bf16x3.ncu-rep --page raw --csv > $out/${tag}_train_bf16x3_raw.csv 2>/dev/null
ls -la $out/*.ncu-rep; rm -f $out/${tag}_train_bf16x3.ncu-rep
[ $(stat -c %s $out/${tag}_fwd_infer_bf16x3.ncu-rep) -gt 30000000 ] && rm -f $out/${tag}_fwd_infer_bf16x3.ncu-rep
du -sh $out
tail -3 $out/${tag}_pytest_gpu.log; tail -2 $out/${tag}_smoke.log; cut -c1-600 $out/${tag}_bench_bf16x3.json; cut -c1-300 $out/${tag}_bench_reference.json
