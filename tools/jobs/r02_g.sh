#!/bin/bash
set -u
O=gpurun_out/r02g; mkdir -p $O
timeout 600 python tools/ab_lib.py build/libklp_14688b5.so build/libklp_mode.so build/libklp_v515.so > $O/ab4.txt 2>&1; tail -4 $O/ab4.txt
timeout 200 python - > $O/phases.txt 2>&1 <<'PY'
import lineprofiles_b200 as klp
from lineprofiles_b200 import synthetic as syn
gt = klp.Table.from_arrays(syn.make_table(), 0)
gt.set_option("phase_clock", 1)
E = syn.energy_grid_linear(1000)
for model, p in (("kline", syn.CONFIG1_PARAMS), ("kline5", syn.random_params("kline5", 1, seed=3))):
    for _ in range(3):
        gt.eval_batch(model, p, E, precision="f32")
    print(model, len(p), {k: round(v, 1) for k, v in gt.phase_times().items()}, gt.last_launch_info())
PY
cat $O/phases.txt
timeout 300 python tools/latency_breakdown.py 2>&1 | grep "exported\|coop=1 B=    1"
