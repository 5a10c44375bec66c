#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 > $out/r2g_pytest_gpu.log 2>&1; echo "rc=$?" >> $out/r2g_pytest_gpu.log
timeout 300 python tools/kernel_times.py bf16x3 > $out/r2g_ktimes_x3.txt 2>&1
for n in 128 112 96; do SCNERF_WGRAD_CTAS=$n timeout 300 python tools/kernel_times.py bf16x3 > $out/r2g_ktimes_x3_wgrad$n.txt 2>&1; done
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > $out/r2g_bench_c2.json 2> $out/r2g_bench_c2.err
timeout 600 python bench.py --steps 20 --warmup 3 --workload c3 --no-cpu-baseline > $out/r2g_bench_c3.json 2> $out/r2g_bench_c3.err
timeout 300 python tools/c3_breakdown.py > $out/r2g_c3_breakdown.txt 2>&1
grep -v Warning $out/r2g_pytest_gpu.log | tail -n 8 | cut -c1-300
for f in $out/r2g_ktimes_x3*.txt; do echo $f; grep -v Warn $f | grep -E "sum of|fwd_pipe|wgrad_kernel|dgrad_pipe|head_wgrad"; done
tail -n 11 $out/r2g_c3_breakdown.txt
cut -c1-260 $out/r2g_bench_c2.json; echo; cut -c1-260 $out/r2g_bench_c3.json; tail -n 3 $out/r2g_bench_c2.err $out/r2g_bench_c3.err
