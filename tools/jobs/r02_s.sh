#!/bin/bash
# round 2, GPU job S (1 GPU): short final staging chunk of the host-buffer pipeline (stage_tail) against slot size
set -u
O=gpurun_out/r02s; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "host_pipeline or device_entry or xspec" > $O/pytest.log 2>&1; tail -3 $O/pytest.log
timeout 600 python tools/e2e_probe.py 0,1024,2048,4096 > $O/e2e.txt 2>&1; cat $O/e2e.txt
