#!/bin/bash
# round 2, GPU job D (1 GPU): full GPU test suite (cooperative small-batch path, new conv kernels), latency breakdown,
# k_quad CTA-size probe incl. 768 threads x 85 registers, ncu of the new conv kernels
set -u
O=gpurun_out/r02d; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log; tail -6 $O/pytest.log
python tools/latency_breakdown.py > $O/latency_breakdown.txt 2>&1; cat $O/latency_breakdown.txt
python - > $O/nt768.txt 2>&1 <<'PY'
import os, sys
sys.path.insert(0, "tools")
from gpu_perf import run
import lineprofiles_b200 as klp
from lineprofiles_b200 import synthetic as syn
gt = klp.Table.from_arrays(syn.make_table(), 0)
for B in (2000, 50000):
    for nt in (1024, 768, 512, 1024, 768):
        run(gt, "kline5", B, "f32", opts={"quad_threads": nt}, reps=2)
PY
cat $O/nt768.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_conv_exact -s 1 -c 1 -o $O/conv_exact2 -f \
  python tools/prof_one.py kconv f32 512 4096 > $O/ncu_conv.log 2>&1; echo "ncu conv exit $?"
KLP_OPTIONS=conv_mode=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_conv_cum -s 1 -c 1 -o $O/conv_cum2 -f \
  python tools/prof_one.py kconv f32 512 4096 > $O/ncu_convf.log 2>&1; echo "ncu convcum exit $?"
