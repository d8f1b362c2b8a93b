#!/bin/bash
# round-2 experiment I: k_model flat path (bin-parallel coding of the samples left after zero runs and chains)
mkdir -p gpurun_out
L=gpurun_out/r2i.log
: > $L
(timeout 900 python -m pytest tests/test_ffv1_gpu.py -m gpu -x -q 2>&1 | tail -4) >> $L
run() { echo "== B=${B:-128} ${K:-grain} $*" >> $L; env "$@" python tools/probe_content.py ${B:-128} ${K:-grain} 2>&1 | grep "^B=\|kernel" | tail -2 >> $L; }
PROBE_KERNELS=1 run X=1
B=64 run X=1
K=flat PROBE_KERNELS=1 run X=1
K=white B=32 PROBE_KERNELS=1 run X=1
K=zero B=32 PROBE_KERNELS=1 run X=1
PROBE_W=2048 PROBE_H=1556 PROBE_LAYOUT=2 PROBE_SLICES=4 K=grain PROBE_KERNELS=1 run X=1
echo "== trace" >> $L
B200_TRACE=1 python tools/probe_content.py 128 grain 2>&1 | grep -A8 "^band" | head -10 >> $L
cat $L
