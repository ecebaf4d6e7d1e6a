#!/bin/bash
# round 2, GPU job W (1 GPU): deposit ring with consumed lines discarded from the L2 (ring_discard): parity first (under timeout), then
# throughput and DRAM traffic with and without
set -u
O=gpurun_out/r02w; mkdir -p $O
timeout 240 python - > $O/ring_smoke.txt 2>&1 <<'PY'
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
        for disc in (0, 1):
            gt.set_option("deposit_ring", 1); gt.set_option("ring_discard", disc)
            b = gt.eval_batch(model, p, Eg, spectrum=spec, precision=prec)
            print(model, prec, "discard", disc, "ring == planes:", np.array_equal(a, b, equal_nan=True), gt.last_launch_info(), flush=True)
# repeated large batches: a discarded line that is reused must never be read before it is rewritten
gt.set_option("deposit_ring", 0)
p = syn.random_params("kline5", 20000, seed=9)
a = gt.eval_batch("kline5", p, E, precision="f32")
gt.set_option("deposit_ring", 1); gt.set_option("ring_discard", 1)
for rep in range(3):
    b = gt.eval_batch("kline5", p, E, precision="f32")
    print("20000 sets, repetition", rep, "ring+discard == planes:", np.array_equal(a, b), flush=True)
PY
RS=$?; cat $O/ring_smoke.txt; echo "ring smoke exit $RS"
if [ $RS -ne 0 ] || grep -q "False" $O/ring_smoke.txt; then echo "RING BROKEN: stopping here"; exit 1; fi
for disc in 0 1; do
  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-configs --option deposit_ring=1 --option ring_discard=$disc | grep '^{' > $O/bench_ring_discard$disc.json
  python -c "import json; d=json.load(open('$O/bench_ring_discard$disc.json')); print('ring discard $disc', round(d['value']), d['ms_per_step'], d['roofline']['kernel'])"
  timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:k_quad -s 3 -c 1 --csv --log-file $O/ring_dram$disc.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-configs --option deposit_ring=1 --option ring_discard=$disc > $O/ncu_ring$disc.log 2>&1; echo "ncu ring exit $?"; tail -4 $O/ring_dram$disc.csv | rev | cut -d, -f1-3 | rev
done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "ring or invariance" > $O/pytest.log 2>&1; tail -3 $O/pytest.log
