#!/bin/bash
# round 2, GPU job Z (1 GPU): final check of the build that ships -- smoke, full GPU suite, bench line + reference arm, ncu of k_conv_exact
# (the v3 capture still executed both index computations)
set -u
O=gpurun_out/r02z; mkdir -p $O
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; echo "smoke exit $?"; tail -1 $O/smoke.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log; tail -4 $O/pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 2> $O/bench.err | grep '^{' > $O/bench.json; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>> $O/bench.err | grep '^{' > $O/bench_reference.json; echo "reference arm exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_conv_exact -s 1 -c 1 -o $O/conv_exact -f \
  python tools/prof_one.py kconv f32 512 4096 > $O/ncu_conv.log 2>&1; echo "ncu conv exit $?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02z/bench.json"))
print("bench", round(d["value"]), d["ms_per_step"], "e2e", round(d["e2e"]["value"]), d["roofline"]["kernel"], d["roofline"]["avg_launch_ms"])
print("issue", d.get("roofline_issue", {}).get("frac"), "fp32", d.get("roofline_fp32_pipe", {}).get("frac"), "compute", d.get("roofline_compute", {}).get("frac"), "traffic", d["roofline"].get("traffic"))
c = d["configs"]["config3_kconv_4096"]; print("config3", c["reference_order"]["spectra_per_s"], c["cumulative_profile"]["spectra_per_s"])
r = json.load(open("gpurun_out/r02z/bench_reference.json")); print("reference arm", r["value"], r.get("cpu_baseline", {}).get("cores"))
PY
