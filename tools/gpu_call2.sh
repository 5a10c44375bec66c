#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x > $out/r2b_pytest_gpu.log 2>&1; echo "rc=$?" >> $out/r2b_pytest_gpu.log
timeout 900 python -m pytest tests/test_gpu_parity_pp.py tests/test_gpu_parity.py -m gpu -q --timeout 900 -s -k "cascade_64_128 or c3_composed or full_size_step_vs_oracle or subpixel" > $out/r2b_pytest_new.log 2>&1; echo "rc=$?" >> $out/r2b_pytest_new.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/r2b_smoke.log 2>&1; echo "rc=$?" >> $out/r2b_smoke.log
tail -n 30 $out/r2b_pytest_gpu.log; grep -v Warning $out/r2b_pytest_new.log | tail -n 60; tail -n 5 $out/r2b_smoke.log
