#!/bin/bash
# round 2, GPU job T (1 GPU): bench line + reference arm + full GPU suite of the final build (capture JSON with the build's fingerprint in place)
set -u
O=gpurun_out/r02t; mkdir -p $O
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; echo "smoke exit $?"; tail -1 $O/smoke.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log; tail -4 $O/pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 2> $O/bench.err | grep '^{' > $O/bench.json; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>> $O/bench.err | grep '^{' > $O/bench_reference.json; echo "reference arm exit $?"
timeout 300 python tools/gpu_perf.py 20000 > $O/perf.txt 2>&1; tail -12 $O/perf.txt
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02t/bench.json"))
print("bench", round(d["value"]), d["ms_per_step"], "e2e", round(d["e2e"]["value"]), d["roofline"]["kernel"], d["roofline"]["avg_launch_ms"])
print("issue", d.get("roofline_issue", {}).get("frac"), "fp32", d.get("roofline_fp32_pipe", {}).get("frac"), "compute", d.get("roofline_compute", {}).get("frac"), "traffic", d["roofline"].get("traffic"))
r = json.load(open("gpurun_out/r02t/bench_reference.json")); print("reference arm", r["value"], r.get("cpu_baseline", {}).get("cores"))
PY
