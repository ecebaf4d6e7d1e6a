#!/bin/bash
# round 2, GPU job E: conv kernels after software pipelining, small-batch phase stamps
set -u
O=gpurun_out/r02e; mkdir -p $O
python tools/conv_perf.py 5000 > $O/conv_perf.txt 2>&1; echo "conv_perf exit $?"; cat $O/conv_perf.txt
python - > $O/phases.txt 2>&1 <<'PY'
import sys, time
import numpy as np, torch
import lineprofiles_b200 as klp
from lineprofiles_b200 import synthetic as syn
gt = klp.Table.from_arrays(syn.make_table(), 0)
gt.set_option("phase_clock", 1)
E = syn.energy_grid_linear(1000)
for model, p in (("kline", syn.CONFIG1_PARAMS), ("kline5", syn.random_params("kline5", 1, seed=3)), ("kline5", syn.random_params("kline5", 8, seed=3))):
    for _ in range(3):
        gt.eval_batch(model, p, E, precision="f32")
    ph = gt.phase_times()
    print(model, len(p), {k: round(v, 1) for k, v in ph.items()}, gt.last_launch_info())
gt.set_option("phase_clock", 0)
PY
cat $O/phases.txt
python tools/latency_breakdown.py > $O/latency_breakdown.txt 2>&1; grep "B=    1\|B=    8\|exported" $O/latency_breakdown.txt
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "coop or conv or irregular or config4" 2>&1 | tail -3
