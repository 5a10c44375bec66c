#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 > $out/r2d_pytest_gpu.log 2>&1; echo "rc=$?" >> $out/r2d_pytest_gpu.log
timeout 600 python bench.py --steps 30 --warmup 5 > $out/r2d_bench_c2.json 2> $out/r2d_bench_c2.err
timeout 600 python bench.py --steps 20 --warmup 3 --workload c3 --no-cpu-baseline > $out/r2d_bench_c3.json 2> $out/r2d_bench_c3.err
timeout 600 python bench.py --steps 20 --warmup 3 --workload c4 --no-cpu-baseline > $out/r2d_bench_c4.json 2> $out/r2d_bench_c4.err
timeout 600 python bench.py --steps 10 --warmup 3 --workload c5 --no-cpu-baseline > $out/r2d_bench_c5.json 2> $out/r2d_bench_c5.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/r2d_bench_ref_c2.json 2> $out/r2d_bench_ref_c2.err
grep -v Warning $out/r2d_pytest_gpu.log | tail -n 40 | cut -c1-300
for w in c2 c3 c4 c5 ref_c2; do echo "== $w"; cut -c1-1500 $out/r2d_bench_$w.json; tail -n 5 $out/r2d_bench_$w.err; done
