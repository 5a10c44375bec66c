#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests -m gpu -x -q --timeout 600 > $out/r2a_pytest_gpu.log 2>&1; echo "rc=$?" >> $out/r2a_pytest_gpu.log
for r in 0 1 2; do
  SCNERF_EPI_ROLL=$r timeout 300 python tools/kernel_times.py bf16x3 > $out/r2a_ktimes_x3_roll$r.txt 2>&1
done
for r in 1 2; do
  SCNERF_EPI_ROLL=$r timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "bf16x3 or tc or full_size" --timeout 600 > $out/r2a_pytest_roll$r.log 2>&1; echo "rc=$?" >> $out/r2a_pytest_roll$r.log
done
timeout 300 python tools/step_timeline.py bf16x3 > $out/r2a_step_timeline_x3.txt 2>&1
for r in 0 1 2; do SCNERF_EPI_ROLL=$r timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $out/r2a_bench_roll$r.json 2>$out/r2a_bench_roll$r.err; done
tail -3 $out/r2a_pytest_gpu.log; head -5 $out/r2a_ktimes_x3_roll*.txt; tail -2 $out/r2a_pytest_roll*.log; tail -3 $out/r2a_step_timeline_x3.txt; cut -c1-200 $out/r2a_bench_roll*.json
