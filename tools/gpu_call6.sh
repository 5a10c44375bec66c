#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 > $out/r2f_pytest_gpu.log 2>&1; echo "rc=$?" >> $out/r2f_pytest_gpu.log
timeout 300 python tools/kernel_times.py bf16x3 > $out/r2f_ktimes_x3.txt 2>&1
timeout 300 python tools/c3_breakdown.py > $out/r2f_c3_breakdown.txt 2>&1
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > $out/r2f_bench_c2.json 2> $out/r2f_bench_c2.err
timeout 600 python bench.py --steps 20 --warmup 3 --workload c3 --no-cpu-baseline > $out/r2f_bench_c3.json 2> $out/r2f_bench_c3.err
grep -v Warning $out/r2f_pytest_gpu.log | tail -n 12 | cut -c1-300
grep -v Warn $out/r2f_ktimes_x3.txt | tail -n 9; tail -n 12 $out/r2f_c3_breakdown.txt
cut -c1-260 $out/r2f_bench_c2.json; echo; cut -c1-260 $out/r2f_bench_c3.json; tail -n 3 $out/r2f_bench_c2.err $out/r2f_bench_c3.err
