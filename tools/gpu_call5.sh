#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 > $out/r2e_pytest_gpu.log 2>&1; echo "rc=$?" >> $out/r2e_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 --workload c3 --no-cpu-baseline > $out/r2e_bench_c3.json 2> $out/r2e_bench_c3.err
grep -v Warning $out/r2e_pytest_gpu.log | tail -n 40 | cut -c1-400
cut -c1-400 $out/r2e_bench_c3.json; tail -n 5 $out/r2e_bench_c3.err
