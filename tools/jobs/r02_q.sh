#!/bin/bash
# round 2, GPU job Q (1 GPU): evidence for the build that ships after the one-correction division -- full GPU test suite, bench line
# (+ reference arm), launch list, ncu --set full of k_quad at bench scale and of both convolution kernels, latency breakdown, fuzz
set -u
O=gpurun_out/r02q; mkdir -p $O
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; echo "smoke exit $?"; tail -1 $O/smoke.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log; tail -4 $O/pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 2> $O/bench.err | grep '^{' > $O/bench.json; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>> $O/bench.err | grep '^{' > $O/bench_reference.json; echo "reference arm exit $?"
timeout 300 python tools/latency_breakdown.py > $O/latency_breakdown.txt 2>&1; grep "coop=1 B=    1\|exported" $O/latency_breakdown.txt
timeout 300 python tools/fuzz.py 3 2048 > $O/fuzz.txt 2>&1; tail -4 $O/fuzz.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-configs > $O/ncu_list.log 2>&1; echo "ncu list exit $?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_quad -s 3 -c 1 -o $O/quad_bench -f \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-configs > $O/ncu_quad.log 2>&1; echo "ncu quad exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_conv_exact -s 1 -c 1 -o $O/conv_exact -f \
  python tools/prof_one.py kconv f32 512 4096 > $O/ncu_conv.log 2>&1; echo "ncu conv exit $?"
KLP_OPTIONS=conv_mode=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_conv_cum -s 1 -c 1 -o $O/conv_cum -f \
  python tools/prof_one.py kconv f32 512 4096 > $O/ncu_convf.log 2>&1; echo "ncu convcum exit $?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02q/bench.json"))
print("bench", round(d["value"]), d["ms_per_step"], "e2e", round(d["e2e"]["value"]), d["roofline"]["kernel"], d["roofline"]["avg_launch_ms"])
PY
ls -la $O | head -30
