#!/bin/bash
# round 2, GPU job AC (1 GPU): exact zeros in the table's branches stay inside the branch-free arithmetic (per-set check on the blended
# frame): parity on zero-row tables, A/B of the plain mode against the shipped build (the hot loop was re-scheduled), zero-row bench
set -u
O=gpurun_out/r02ac; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "zeroed_rows or fast_and_strict or random_sets or edge_parameters or big_table" > $O/pytest.log 2>&1; tail -3 $O/pytest.log
timeout 300 python tools/ab_lib.py build/libklp_ship.so lineprofiles_b200/libklineprofiles_b200.so 2>&1 | tail -2
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e | grep '^{' > $O/bench.json
python -c "import json; d=json.load(open('$O/bench.json')); print(round(d['value']), d['configs']['table_with_zero_rows'])"
