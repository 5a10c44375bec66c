#!/bin/bash
# Final verification on one B200: the driver's own sequence (GPU tests, smoke, reference arm, default bench).
out=gpurun_out; mkdir -p $out; rm -f $out/r2_parity_counts.json
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 > $out/final_pytest_gpu.log 2>&1; echo "rc=$?" >> $out/final_pytest_gpu.log
grep -v Warning $out/final_pytest_gpu.log | tail -n 5 | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/final_smoke.log 2>&1; echo "rc=$?" >> $out/final_smoke.log; tail -n 2 $out/final_smoke.log
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $out/final_bench_reference.json 2> $out/final_bench_reference.err ) 2>&1 | grep real
( time timeout 900 python bench.py > $out/final_bench_default.json 2> $out/final_bench_default.err ) 2>&1 | grep real
python - <<'PY'
import json
for f in ("gpurun_out/final_bench_reference.json", "gpurun_out/final_bench_default.json"):
    d = json.loads([l for l in open(f) if l.startswith("{")][-1])
    print(f, round(d["ms_per_step"], 3), round(d["value"], 1), round(d["e2e"]["value"], 1), d.get("gpu_launches"), (d.get("inference") or {}).get("value"),
          (d.get("cpu_baseline") or {}).get("value"), (d.get("clocks") or {}).get("sm_mhz"), (d.get("roofline") or {}).get("frac"))
PY
tail -n 2 $out/final_bench_default.err
