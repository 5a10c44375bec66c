#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/r2p_smoke.log 2>&1; echo "rc=$?" >> $out/r2p_smoke.log
tail -n 4 $out/r2p_smoke.log
( time timeout 900 python bench.py > $out/r2p_bench_default.json 2> $out/r2p_bench_default.err ) 2>&1 | grep real
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r2p_bench_default.json") if l.startswith("{")][-1])
print({k: d[k] for k in ("value", "ms_per_step", "steps", "gpu_launches")}, d["e2e"]["value"], d.get("inference"), d.get("cpu_baseline"), d["clocks"])
print([(k["kernel"], round(k["ms_per_step"], 3)) for k in d["roofline"]["kernels"]], d["roofline"]["bound"], d["roofline"]["frac"])
PY
tail -n 3 $out/r2p_bench_default.err
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $out/r2p_bench_reference.json 2> $out/r2p_bench_reference.err ) 2>&1 | grep real
cut -c1-400 $out/r2p_bench_reference.json
