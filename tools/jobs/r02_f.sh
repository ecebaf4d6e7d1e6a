#!/bin/bash
# round 2, GPU job F: consumer-warp deposit ring (parity, throughput, DRAM traffic), small-batch phases after the plan rewrite
set -u
O=gpurun_out/r02f; mkdir -p $O
# the consumer-warp ring first, alone and under a short timeout: a protocol error there would hang, not fail
timeout 120 python - > $O/ring_smoke.txt 2>&1 <<'PY'
import numpy as np
import lineprofiles_b200 as klp
from lineprofiles_b200 import synthetic as syn
gt = klp.Table.from_arrays(syn.make_table(), 0)
E = syn.energy_grid_linear(1000)
for model, B in (("kline5", 700), ("kline", 2000), ("kconv10", 400)):
    Eg = syn.energy_grid_log(512) if model.startswith("kconv") else E
    spec = syn.powerlaw_continuum(Eg) if model.startswith("kconv") else None
    p = syn.random_params(model, B, seed=5)
    p[:5, 3 if not model.startswith("kconv") else 2] = 1.0    # rmin inside the frame's innermost row: leading radii without candidates
    for prec in ("f32", "f64"):
        gt.set_option("deposit_ring", 0)
        a = gt.eval_batch(model, p, Eg, spectrum=spec, precision=prec)
        gt.set_option("deposit_ring", 1)
        b = gt.eval_batch(model, p, Eg, spectrum=spec, precision=prec)
        print(model, prec, "ring == planes:", np.array_equal(a, b, equal_nan=True), gt.last_launch_info(), flush=True)
PY
RS=$?; cat $O/ring_smoke.txt; echo "ring smoke exit $RS"
if [ $RS -ne 0 ] || grep -q "False" $O/ring_smoke.txt; then echo "RING BROKEN: stopping here"; exit 1; fi
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log; tail -6 $O/pytest.log
timeout 120 python - > $O/phases.txt 2>&1 <<'PY'
import lineprofiles_b200 as klp
from lineprofiles_b200 import synthetic as syn
gt = klp.Table.from_arrays(syn.make_table(), 0)
gt.set_option("phase_clock", 1)
E = syn.energy_grid_linear(1000)
for model, p in (("kline", syn.CONFIG1_PARAMS), ("kline5", syn.random_params("kline5", 1, seed=3)), ("kline5", syn.random_params("kline5", 8, seed=3))):
    for _ in range(3):
        gt.eval_batch(model, p, E, precision="f32")
    print(model, len(p), {k: round(v, 1) for k, v in gt.phase_times().items()}, gt.last_launch_info())
PY
cat $O/phases.txt
timeout 300 python tools/latency_breakdown.py > $O/latency_breakdown.txt 2>&1; grep "B=    1\|B=    8\|exported" $O/latency_breakdown.txt
for ring in 0 1; do
  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-configs --option deposit_ring=$ring | grep '^{' > $O/bench_ring$ring.json
  python -c "import json; d=json.load(open('$O/bench_ring$ring.json')); print('ring $ring', round(d['value']), d['roofline']['kernel'])"
done
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:k_quad -s 6 -c 1 --csv --log-file $O/ring_dram.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-configs --option deposit_ring=1 > $O/ncu_ring.log 2>&1; echo "ncu ring exit $?"; tail -5 $O/ring_dram.csv
