#!/bin/bash
set -u
O=gpurun_out/r02l; mkdir -p $O
timeout 180 python - > $O/ring_smoke.txt 2>&1 <<'PY'
import numpy as np
import lineprofiles_b200 as klp
from lineprofiles_b200 import synthetic as syn
gt = klp.Table.from_arrays(syn.make_table(), 0)
E = syn.energy_grid_linear(1000)
for model, B in (("kline5", 700), ("kline", 2000), ("kconv10", 400)):
    Eg = syn.energy_grid_log(512) if model.startswith("kconv") else E
    spec = syn.powerlaw_continuum(Eg) if model.startswith("kconv") else None
    p = syn.random_params(model, B, seed=5)
    p[:5, 3 if not model.startswith("kconv") else 2] = 1.0
    for prec in ("f32", "f64"):
        gt.set_option("deposit_ring", 0)
        a = gt.eval_batch(model, p, Eg, spectrum=spec, precision=prec)
        gt.set_option("deposit_ring", 1)
        b = gt.eval_batch(model, p, Eg, spectrum=spec, precision=prec)
        print(model, prec, "ring == planes:", np.array_equal(a, b, equal_nan=True), gt.last_launch_info(), flush=True)
PY
RS=$?; cat $O/ring_smoke.txt; echo "ring smoke exit $RS"
if [ $RS -ne 0 ] || grep -q "False" $O/ring_smoke.txt; then echo "RING BROKEN: stopping here"; exit 1; fi
timeout 300 python tools/ab_lib.py build/libklp_14688b5.so build/libklp_mode.so lineprofiles_b200/libklineprofiles_b200.so 2>&1 | tail -3
for ring in 0 1; do
  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-configs --option deposit_ring=$ring | grep '^{' > $O/bench_ring$ring.json
  python -c "import json; d=json.load(open('$O/bench_ring$ring.json')); print('ring $ring', round(d['value']), d['ms_per_step'], d['roofline']['kernel'])"
done
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:k_quad -s 3 -c 1 --csv --log-file $O/ring_dram.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-configs --option deposit_ring=1 > $O/ncu_ring.log 2>&1; echo "ncu ring exit $?"; tail -4 $O/ring_dram.csv | cut -d, -f13-15
