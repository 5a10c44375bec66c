#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_parity_pp.py -m gpu -q -x --timeout 600 > $out/r2o_pytest_pp.log 2>&1; echo "rc=$?" >> $out/r2o_pytest_pp.log
grep -v Warning $out/r2o_pytest_pp.log | tail -n 15 | cut -c1-300
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > $out/r2o_pytest_gpu.log 2>&1; echo "rc=$?" >> $out/r2o_pytest_gpu.log
grep -v Warning $out/r2o_pytest_gpu.log | tail -n 8 | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 3 --workload c4 --no-cpu-baseline > $out/r2o_bench_c4.json 2> $out/r2o_bench_c4.err
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > $out/r2o_bench_c2.json 2> $out/r2o_bench_c2.err
cut -c1-200 $out/r2o_bench_c4.json; echo; cut -c1-200 $out/r2o_bench_c2.json; tail -n 3 $out/r2o_bench_c4.err $out/r2o_bench_c2.err
python - <<'PY'
import json
for f in ("gpurun_out/r2o_bench_c4.json", "gpurun_out/r2o_bench_c2.json"):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, [(k["kernel"], round(k["ms_per_step"], 3)) for k in d["roofline"]["kernels"]], d["roofline"].get("other_kernels_ms_per_step"))
    except Exception as e:
        print(f, "ERR", e)
PY
