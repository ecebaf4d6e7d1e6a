#!/bin/bash
# round 2, GPU job H (8 GPUs): group tests, bench at N = 8 (one process per GPU, and one process driving all eight), N = 1 on the
# same box, config 4 (kconv10, 1e6 walkers, every GPU ends up with all 32.8 GB of spectra)
set -u
O=gpurun_out/r02h; mkdir -p $O
nvidia-smi -L | head -8
timeout 300 python -m pytest tests/test_group.py -m gpu -x -q > $O/pytest_group.log 2>&1; echo "pytest exit $?" >> $O/pytest_group.log; tail -3 $O/pytest_group.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 python bench.py --gpus 1 --steps 3 --warmup 3 --no-configs --no-cpu-baseline 2> $O/b1.err | grep '^{' > $O/bench_1gpu.json
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL timeout 400 $TR bench.py --gpus 8 --steps 3 --warmup 3 2> $O/b8.err | grep '^{' > $O/bench_8gpu_torchrun.json; echo "torchrun exit $?"
grep -i "nvls\|NCCL version\|comm 0x.*nranks" $O/b8.err | head -6
timeout 400 python bench.py --gpus 8 --steps 3 --warmup 3 --no-e2e 2> $O/b8p.err | grep '^{' > $O/bench_8gpu_oneprocess.json; echo "one-process exit $?"
timeout 400 $TR bench.py --gpus 8 --steps 3 --warmup 3 --no-e2e --gather-tail 0 2>> $O/b8.err | grep '^{' > $O/bench_8gpu_torchrun_notail.json
timeout 600 $TR tools/bench_configs.py --configs 4 --no-cpu 2> $O/c4.err | grep '^{' > $O/config4_8gpu.jsonl; echo "config4 exit $?"; cat $O/config4_8gpu.jsonl
python - <<'PY'
import json
O = "gpurun_out/r02h/"
one = json.load(open(O + "bench_1gpu.json"))
print("1 GPU", round(one["value"]), one["ms_per_step"])
for f in ("bench_8gpu_torchrun.json", "bench_8gpu_oneprocess.json", "bench_8gpu_torchrun_notail.json"):
    try:
        d = json.load(open(O + f)); print(f, round(d["value"]), d["ms_per_step"], "eff", round(d["value"] / (8 * one["value"]), 4), d["e2e"] and round(d["e2e"]["value"]), d["config"]["gather"][:70])
    except Exception as e:
        print(f, "ERR", e)
PY
tail -2 $O/*.err | head -40
