#!/bin/bash
# GPU run of the decoder: parity tests, then throughput by slices per warp
mkdir -p gpurun_out
L=gpurun_out/dec1.log
(timeout 900 python -m pytest tests/test_ffv1_dec_gpu.py -m gpu -x -q 2>&1 | tail -25) > $L
(timeout 600 python tools/probe_decode.py 32 grain 1 2 4 2>&1 | tail -6) >> $L
(timeout 600 python tools/probe_decode.py 128 grain 1 2 4 8 2>&1 | tail -6) >> $L
cat $L
