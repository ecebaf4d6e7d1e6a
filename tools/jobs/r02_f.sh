#!/bin/bash
# round 2, GPU job F: consumer-warp deposit ring (parity, throughput, DRAM traffic), small-batch phases after the plan rewrite
set -u
O=gpurun_out/r02f; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log; tail -6 $O/pytest.log
python - > $O/phases.txt 2>&1 <<'PY'
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
python tools/latency_breakdown.py > $O/latency_breakdown.txt 2>&1; grep "B=    1\|B=    8\|exported" $O/latency_breakdown.txt
for ring in 0 1; do
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-configs --option deposit_ring=$ring | grep '^{' > $O/bench_ring$ring.json
  python -c "import json; d=json.load(open('$O/bench_ring$ring.json')); print('ring $ring', round(d['value']), d['roofline']['kernel'])"
done
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:k_quad -s 6 -c 1 --csv --log-file $O/ring_dram.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-configs --option deposit_ring=1 > $O/ncu_ring.log 2>&1; echo "ncu ring exit $?"; tail -5 $O/ring_dram.csv
