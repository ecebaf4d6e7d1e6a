#!/bin/bash
# round 2, GPU job U (8 GPUs): scaling of the build with the one-correction division
set -u
O=gpurun_out/r02u; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 python bench.py --gpus 1 --steps 3 --warmup 3 --no-configs --no-cpu-baseline --no-e2e 2> $O/b1.err | grep '^{' > $O/bench_1gpu.json
timeout 400 $TR bench.py --gpus 8 --steps 3 --warmup 3 2> $O/b8.err | grep '^{' > $O/bench_8gpu_torchrun.json; echo "torchrun exit $?"
timeout 400 python bench.py --gpus 8 --steps 3 --warmup 3 --no-e2e 2> $O/b8p.err | grep '^{' > $O/bench_8gpu_oneprocess.json; echo "one-process exit $?"
python - <<'PY'
import json
O = "gpurun_out/r02u/"
one = json.load(open(O + "bench_1gpu.json"))
print("1 GPU", round(one["value"]), one["ms_per_step"])
for f in ("bench_8gpu_torchrun.json", "bench_8gpu_oneprocess.json"):
    d = json.load(open(O + f)); print(f, round(d["value"]), d["ms_per_step"], "eff", round(d["value"] / (8 * one["value"]), 4), d["e2e"] and round(d["e2e"]["value"]), d["config"]["gather"][:70])
PY
