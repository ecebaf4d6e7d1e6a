#!/bin/bash
# round 2, GPU job A: parity tests, bench line, deposit_ring A/B with 1024-thread CTAs, launch list, ncu --set full of the
# kernels that ship (k_quad<float,1024,1> at bench scale, k_conv, k_conv_fast), single-call latency breakdown.
set -u
O=gpurun_out/r02a; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log; tail -5 $O/pytest.log
python bench.py --steps 3 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench exit $?"; head -c 1500 $O/bench.json
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-configs --option deposit_ring=1 > $O/bench_ring.json 2>> $O/bench.err; echo "ring exit $?"
python tools/latency_breakdown.py > $O/latency_breakdown.txt 2>&1
python tools/latency_probe.py > $O/latency_probe.txt 2>&1
# launch list (serialised, cold caches: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-configs > $O/ncu_list.log 2>&1; echo "ncu list exit $?"
# one bench-scale k_quad launch (the first launch of the timed step: 65535 sets), full set with source
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_quad -s 6 -c 1 -o $O/quad_bench -f \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-configs > $O/ncu_quad.log 2>&1; echo "ncu quad exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_conv -s 1 -c 1 -o $O/conv_exact -f \
  python tools/prof_one.py kconv f32 512 4096 > $O/ncu_conv.log 2>&1; echo "ncu conv exit $?"
KLP_OPTIONS=conv_mode=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_conv_fast -s 1 -c 1 -o $O/conv_fast -f \
  python tools/prof_one.py kconv f32 512 4096 > $O/ncu_convf.log 2>&1; echo "ncu convfast exit $?"
ls -la $O
