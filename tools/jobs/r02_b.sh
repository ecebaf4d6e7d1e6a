#!/bin/bash
# round 2, GPU job B (2 GPUs): the group tests, and the bench at N = 2 both ways (torchrun: one process per GPU; one process
# driving both GPUs), with the gather chunked/overlapped vs a single un-overlapped all-gather (gather_chunk 65535).
set -u
O=gpurun_out/r02b; mkdir -p $O
nvidia-smi -L
python -m pytest tests/test_group.py -m gpu -x -q > $O/pytest_group.log 2>&1; echo "pytest exit $?" >> $O/pytest_group.log; tail -5 $O/pytest_group.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 2 --steps 3 --warmup 3 > $O/bench_2gpu_torchrun.json 2> $O/bench_2gpu_torchrun.err; echo "torchrun exit $?"; head -c 600 $O/bench_2gpu_torchrun.json; echo
$TR bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e --gather-chunk 65535 > $O/bench_2gpu_torchrun_nochunk.json 2>> $O/bench_2gpu_torchrun.err; echo "torchrun nochunk exit $?"
python bench.py --gpus 2 --steps 3 --warmup 3 > $O/bench_2gpu_oneprocess.json 2> $O/bench_2gpu_oneprocess.err; echo "one-process exit $?"; head -c 600 $O/bench_2gpu_oneprocess.json; echo
python bench.py --gpus 1 --steps 3 --warmup 3 --no-configs --no-cpu-baseline > $O/bench_1gpu.json 2> $O/bench_1gpu.err; echo "1gpu exit $?"
tail -3 $O/*.err
