#!/bin/bash
# round 2, GPU job R (1 GPU): k_conv_exact with the conversion-free index as a template parameter (the run-time choice was if-converted)
set -u
O=gpurun_out/r02r; mkdir -p $O
timeout 300 python tools/conv_ab.py > $O/conv_new.txt 2>&1; cat $O/conv_new.txt | tail -12
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "conv or config3 or config4 or ragged or irregular or golden or random_sets" > $O/pytest_conv.log 2>&1; tail -3 $O/pytest_conv.log
