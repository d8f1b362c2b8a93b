#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/scan1.log
(timeout 600 python -m pytest tests/test_scan.py -m gpu -x -q 2>&1 | tail -15) > $L
(timeout 600 python tools/probe_scan.py 128 2>&1 | tail -8) >> $L
(timeout 600 python tools/probe_scan.py 512 2>&1 | grep "^md5") >> $L
cat $L
