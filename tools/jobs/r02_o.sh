#!/bin/bash
# round 2, GPU job O (1 GPU): one-correction Markstein division (3 operations instead of 5) in k_quad and k_conv_exact, conversion-free
# table indices in the convolution kernels, k_rebin shared memory sized to the grid: parity first, then A/B against the previous build
set -u
O=gpurun_out/r02o; mkdir -p $O
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; echo "smoke exit $?"; tail -1 $O/smoke.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log; tail -6 $O/pytest.log
timeout 300 python tools/ab_lib.py build/libklp_5op.so lineprofiles_b200/libklineprofiles_b200.so > $O/ab.txt 2>&1; tail -2 $O/ab.txt
KLP_LIBRARY=$PWD/build/libklp_5op.so timeout 300 python tools/conv_ab.py > $O/conv_old.txt 2>&1; cat $O/conv_old.txt | tail -6
timeout 300 python tools/conv_ab.py > $O/conv_new.txt 2>&1; cat $O/conv_new.txt | tail -6
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2> $O/bench.err | grep '^{' > $O/bench.json; echo "bench exit $?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02o/bench.json"))
print("bench", round(d["value"]), d["ms_per_step"], "e2e", round(d["e2e"]["value"]), d["roofline"]["kernel"], d["roofline"]["avg_launch_ms"])
c = d.get("configs") or {}
print({k: (v.get("ms_per_call") or {m: v[m].get("ms_per_batch") for m in v if isinstance(v[m], dict) and "ms_per_batch" in v[m]} or v.get("fused_kline_spectra_per_s")) for k, v in c.items() if isinstance(v, dict)})
PY
