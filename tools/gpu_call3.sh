#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 > $out/r2c_pytest_gpu.log 2>&1; echo "rc=$?" >> $out/r2c_pytest_gpu.log
timeout 600 python tools/pp_grad_precision.py > $out/r2c_pp_grad_precision_pipe1.txt 2>&1
SCNERF_DGRAD_PIPE=0 timeout 600 python tools/pp_grad_precision.py > $out/r2c_pp_grad_precision_pipe0.txt 2>&1
grep -v Warning $out/r2c_pytest_gpu.log | tail -n 60; cat $out/r2c_pp_grad_precision_pipe1.txt | tail -n 50; tail -n 50 $out/r2c_pp_grad_precision_pipe0.txt
