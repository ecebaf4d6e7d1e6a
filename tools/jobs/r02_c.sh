#!/bin/bash
# round 2, GPU job C (2 GPUs): group tests again, conv kernel A/B, irregular-edge and conv parity tests, 2-GPU bench with the new chunking
set -u
O=gpurun_out/r02c; mkdir -p $O
python -m pytest tests/test_group.py -m gpu -x -q > $O/pytest_group.log 2>&1; echo "pytest group exit $?" >> $O/pytest_group.log; tail -5 $O/pytest_group.log
python tools/conv_perf.py 5000 > $O/conv_perf.txt 2>&1; echo "conv_perf exit $?"; cat $O/conv_perf.txt
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "conv or irregular or ragged or config4 or random_sets or golden or xspec" > $O/pytest_conv.log 2>&1; echo "pytest conv exit $?" >> $O/pytest_conv.log; tail -5 $O/pytest_conv.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e 2> $O/bench_2gpu.err | grep '^{' > $O/bench_2gpu_torchrun.json; echo "torchrun exit $?"
python bench.py --gpus 1 --steps 3 --warmup 3 --no-configs --no-cpu-baseline --no-e2e | grep '^{' > $O/bench_1gpu.json
python - <<'PY'
import json
for f in ("bench_1gpu.json", "bench_2gpu_torchrun.json"):
    d = json.load(open("gpurun_out/r02c/" + f)); print(f, round(d["value"]), d["ms_per_step"], d["config"]["gather"][:80])
PY
