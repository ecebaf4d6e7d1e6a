#!/bin/bash
# round 2, GPU job J (4 GPUs): scaling with the final chunking (one launch of 97952 sets + a 2048-set tail)
set -u
O=gpurun_out/r02j; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 python bench.py --gpus 1 --steps 3 --warmup 3 --no-configs --no-cpu-baseline --no-e2e 2> $O/b1.err | grep '^{' > $O/bench_1gpu.json
for n in 2 4; do
  timeout 400 $TR --nproc-per-node $n bench.py --gpus $n --steps 3 --warmup 3 --no-e2e 2> $O/b$n.err | grep '^{' > $O/bench_${n}gpu.json; echo "N=$n exit $?"
done
timeout 400 $TR --nproc-per-node 4 bench.py --gpus 4 --steps 3 --warmup 3 --no-e2e --gather-tail 512 2>> $O/b4.err | grep '^{' > $O/bench_4gpu_tail512.json
python - <<'PY'
import json
O = "gpurun_out/r02j/"
one = json.load(open(O + "bench_1gpu.json"))
print("1 GPU", round(one["value"]), one["ms_per_step"])
for f in ("bench_2gpu.json", "bench_4gpu.json", "bench_4gpu_tail512.json"):
    d = json.load(open(O + f)); n = d["n_gpus"]
    print(f, round(d["value"]), d["ms_per_step"], "eff", round(d["value"] / (n * one["value"]), 4), d["config"]["gather"][:60], d["roofline"]["avg_launch_ms"], d["roofline"]["launches_timed"])
PY
