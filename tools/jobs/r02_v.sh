#!/bin/bash
# round 2, GPU job V (1 GPU): deposit ring and f64 re-measured on the one-correction build
set -u
O=gpurun_out/r02v; mkdir -p $O
for ring in 0 1; do
  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-configs --option deposit_ring=$ring | grep '^{' > $O/bench_ring$ring.json
  python -c "import json; d=json.load(open('$O/bench_ring$ring.json')); print('ring $ring', round(d['value']), d['ms_per_step'], d['roofline']['kernel'])"
done
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-configs --precision f64 --batch 50000 | grep '^{' > $O/bench_f64.json
python -c "import json; d=json.load(open('$O/bench_f64.json')); print('f64', round(d['value']), d['ms_per_step'], d['roofline']['kernel'])"
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:k_quad -s 3 -c 1 --csv --log-file $O/ring_dram.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-configs --option deposit_ring=1 > $O/ncu_ring.log 2>&1; echo "ncu ring exit $?"; tail -4 $O/ring_dram.csv | cut -d, -f13-15
