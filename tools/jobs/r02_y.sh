#!/bin/bash
# round 2, GPU job Y (8 GPUs): config 4 (kconv10, 1e6 walkers, every GPU ends up with all 32.8 GB of spectra) and the two-GPU group
# tests on the final build
set -u
O=gpurun_out/r02y; mkdir -p $O
timeout 300 python -m pytest tests/test_group.py -m gpu -x -q > $O/pytest_group.log 2>&1; echo "pytest exit $?" >> $O/pytest_group.log; tail -3 $O/pytest_group.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR tools/bench_configs.py --configs 4 --no-cpu 2> $O/c4.err | grep '^{' > $O/config4_8gpu.jsonl; echo "config4 exit $?"; cat $O/config4_8gpu.jsonl | cut -c1-600
