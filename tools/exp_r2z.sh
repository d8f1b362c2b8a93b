#!/bin/bash
mkdir -p gpurun_out
TAG=${TAG:-r2z}
L=gpurun_out/$TAG.log
: > $L
(timeout 900 python -m pytest tests/test_ffv1_gpu.py -m gpu -x -q 2>&1 | tail -2) >> $L
run() { echo "== B=${B:-128} ${K:-grain} $*" >> $L; env "$@" python tools/probe_content.py ${B:-128} ${K:-grain} 2>&1 | grep "^B=\|kernel" | tail -2 >> $L; }
PROBE_KERNELS=1 run X=1
B200_TRACE=1 python tools/probe_content.py 128 grain 2>&1 | grep -A6 "^band" | tail -4 >> $L
B=64 run X=1
B=192 run B200_RANGE_CTAS_PER_SM=2
PROBE_W=2048 PROBE_H=1556 PROBE_LAYOUT=2 PROBE_SLICES=4 K=grain PROBE_KERNELS=1 run X=1
cat $L
