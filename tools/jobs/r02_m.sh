#!/bin/bash
set -u
O=gpurun_out/r02m; mkdir -p $O
timeout 400 python tools/ab_lib.py lineprofiles_b200/libklineprofiles_b200.so build/pads/libklp_pad32.so build/pads/libklp_pad64.so build/pads/libklp_pad96.so > $O/ab_pads.txt 2>&1; tail -5 $O/ab_pads.txt
