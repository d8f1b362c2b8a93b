#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r4e.log
(timeout 900 python -m pytest tests/test_ffv1_dec_gpu.py tests/test_scan.py -m gpu -x -q 2>&1 | tail -4) > $L
(timeout 600 python tools/probe_decode.py 32 grain 1 2>&1 | grep "^decode") >> $L
(timeout 600 python tools/probe_decode.py 128 grain 2 4 2>&1 | grep "^decode") >> $L
(timeout 600 python tools/probe_decode.py 256 grain 3 4 2>&1 | grep "^decode") >> $L
(timeout 600 python tools/probe_decode.py 128 flat 2 2>&1 | grep "^decode") >> $L
(PROBE_W=2048 PROBE_H=1556 PROBE_LAYOUT=2 PROBE_SLICES=4 timeout 600 python tools/probe_decode.py 128 grain 1 2>&1 | grep "^decode") >> $L
(timeout 600 python tools/probe_scan.py 128 2>&1 | grep "^padding") >> $L
cat $L
