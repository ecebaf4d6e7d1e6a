#!/bin/bash
# round 2, GPU job AD (1 GPU): evidence for the build that ships (zeros allowed in the branches; plain-mode loop re-drawn): smoke, full
# GPU suite, ncu launch list and --set full of k_quad at bench scale.  The bench line is taken by job AE, after the capture JSON of
# this build is in place (bench.py refuses ncu figures of another hot loop).
set -u
O=gpurun_out/r02ad; mkdir -p $O
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; echo "smoke exit $?"; tail -1 $O/smoke.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log; tail -4 $O/pytest.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-configs > $O/ncu_list.log 2>&1; echo "ncu list exit $?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_quad -s 3 -c 1 -o $O/quad_bench -f \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-configs > $O/ncu_quad.log 2>&1; echo "ncu quad exit $?"
