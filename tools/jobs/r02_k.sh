#!/bin/bash
set -u
O=gpurun_out/r02k; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log; tail -4 $O/pytest.log
timeout 300 python tools/latency_breakdown.py > $O/latency_breakdown.txt 2>&1; grep "coop=1 B=    1\|exported" $O/latency_breakdown.txt
timeout 200 python tools/plots.py --out $O/plots > $O/plots.txt 2>&1; cat $O/plots.txt
timeout 300 python tools/ab_lib.py build/libklp_14688b5.so lineprofiles_b200/libklineprofiles_b200.so 2>&1 | tail -2
