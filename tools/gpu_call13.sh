#!/bin/bash
# 2-GPU validation (gpurun --gpus 2): the multi-rank test, then the bench at N=2 (weak and strong; c2 and c4) and N=1 on the same box.
out=gpurun_out; mkdir -p $out
nvidia-smi -L > $out/r2m_gpus.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -s --timeout 900 > $out/r2m_pytest_multi.log 2>&1; echo "rc=$?" >> $out/r2m_pytest_multi.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 python bench.py --gpus 1 --steps 50 --warmup 5 --no-cpu-baseline > $out/r2m_bench_c2_n1.json 2> $out/r2m_bench_c2_n1.err
timeout 600 $TR --master-port 29711 bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu-baseline > $out/r2m_bench_c2_n2_weak.json 2> $out/r2m_bench_c2_n2_weak.err
timeout 600 $TR --master-port 29712 bench.py --gpus 2 --steps 50 --warmup 5 --scaling strong --no-cpu-baseline > $out/r2m_bench_c2_n2_strong.json 2> $out/r2m_bench_c2_n2_strong.err
timeout 600 $TR --master-port 29713 bench.py --gpus 2 --steps 20 --warmup 3 --workload c4 --no-cpu-baseline > $out/r2m_bench_c4_n2_weak.json 2> $out/r2m_bench_c4_n2_weak.err
timeout 600 $TR --master-port 29714 bench.py --gpus 2 --steps 20 --warmup 3 --workload c3 --no-cpu-baseline > $out/r2m_bench_c3_n2_weak.json 2> $out/r2m_bench_c3_n2_weak.err
timeout 300 $TR --master-port 29715 bench.py --gpus 2 --steps 1 --warmup 0 --impl reference > $out/r2m_bench_ref_n2.json 2> $out/r2m_bench_ref_n2.err
grep -v Warning $out/r2m_pytest_multi.log | tail -n 12 | cut -c1-300
for f in c2_n1 c2_n2_weak c2_n2_strong c4_n2_weak c3_n2_weak ref_n2; do echo "== $f"; grep '^{' $out/r2m_bench_$f.json | cut -c1-230; grep -v -i warn $out/r2m_bench_$f.err | tail -n 3 | cut -c1-300; done
