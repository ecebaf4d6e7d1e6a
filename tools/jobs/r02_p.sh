#!/bin/bash
# round 2, GPU job P (1 GPU): y/2 by exponent decrement in the square root (variant), k_conv_cum with the per-source pairs in shared memory
set -u
O=gpurun_out/r02p; mkdir -p $O
timeout 300 python tools/ab_lib.py lineprofiles_b200/libklineprofiles_b200.so build/libklp_halfexp.so > $O/ab.txt 2>&1; tail -2 $O/ab.txt
timeout 300 python tools/conv_ab.py > $O/conv_new.txt 2>&1; cat $O/conv_new.txt | tail -12
KLP_LIBRARY=$PWD/build/libklp_halfexp.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "exact_arithmetic or fast_and_strict or random_sets or config1 or golden" > $O/pytest_halfexp.log 2>&1; tail -3 $O/pytest_halfexp.log
