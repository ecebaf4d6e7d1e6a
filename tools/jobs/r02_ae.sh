#!/bin/bash
# round 2, GPU job AE (1 GPU): the bench line and the reference arm of the build that ships
set -u
O=gpurun_out/r02ae; mkdir -p $O
timeout 300 python bench.py --steps 5 --warmup 3 2> $O/bench.err | grep '^{' > $O/bench.json; echo "bench exit $?"
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 2>> $O/bench.err | grep '^{' > $O/bench_reference.json; echo "reference arm exit $?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02ae/bench.json"))
print("bench", round(d["value"]), d["ms_per_step"], "e2e", round(d["e2e"]["value"]), d["roofline"]["kernel"], d["roofline"]["avg_launch_ms"])
print("issue", (d.get("roofline_issue") or {}).get("frac"), "fp32", (d.get("roofline_fp32_pipe") or {}).get("frac"), "compute", d.get("roofline_compute", {}).get("frac"), "traffic", d["roofline"].get("traffic"), d["roofline"].get("traffic_note", "")[:80])
c = d["configs"]; print("config1", c["config1_kline_single_call"]["ms_per_call"], "config3", c["config3_kconv_4096"]["reference_order"]["spectra_per_s"], c["config3_kconv_4096"]["cumulative_profile"]["spectra_per_s"], "zero rows", c["table_with_zero_rows"]["kline5_spectra_per_s"])
r = json.load(open("gpurun_out/r02ae/bench_reference.json")); print("reference arm", r["value"], r.get("cpu_baseline", {}).get("cores"))
PY
