#!/bin/bash
# round-2 experiment E: k_range with per-lane TMA staging; new CLI pipeline
mkdir -p gpurun_out
L=gpurun_out/r2e.log
: > $L
run() { echo "== B=${B:-128} $*" >> $L; env "$@" python tools/probe_content.py ${B:-128} grain 2>&1 | grep "^B=" | tail -1 >> $L; }
run B200_RANGE_SMS=24
run B200_RANGE_SMS=16
run B200_RANGE_SMS=32
B=64 run B200_RANGE_SMS=24
B=64 run B200_RANGE_SMS=16
B=192 run B200_RANGE_SMS=24
echo "== kernels serial" >> $L
PROBE_KERNELS=1 python tools/probe_content.py 128 grain 2>&1 | grep kernel >> $L
echo "== trace R=24" >> $L
B200_TRACE=1 python tools/probe_content.py 128 grain 2>&1 | grep -A12 "^band" | head -14 >> $L
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12) >> $L
echo "== cli throughput 256 files, 64 in flight" >> $L
timeout 600 python tools/cli_throughput.py 256 >> $L 2>&1
echo "== cli throughput 256 files, 128 in flight" >> $L
B200_FRAMES_IN_FLIGHT=128 timeout 600 python tools/cli_throughput.py 256 >> $L 2>&1
cat $L
