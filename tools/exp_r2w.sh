#!/bin/bash
# k_range occupancy: warps per scheduler vs time per band, B = 128 and 256
mkdir -p gpurun_out
L=gpurun_out/r2w.log
: > $L
run() { echo "== B=${B:-128} ${K:-grain} $*" >> $L; env "$@" python tools/probe_content.py ${B:-128} ${K:-grain} 2>&1 | grep "^B=\|kernel" | tail -2 >> $L; }
PROBE_KERNELS=1 run B200_RANGE_SMS=24
run B200_RANGE_SMS=12 B200_RANGE_CTAS_PER_SM=2
run B200_RANGE_SMS=16 B200_RANGE_CTAS_PER_SM=2
B=256 PROBE_KERNELS=1 run B200_RANGE_SMS=24 B200_RANGE_CTAS_PER_SM=2
B=192 run B200_RANGE_SMS=24 B200_RANGE_CTAS_PER_SM=2
B=192 run B200_RANGE_SMS=20 B200_RANGE_CTAS_PER_SM=2
echo "== trace B=256" >> $L
B200_RANGE_CTAS_PER_SM=2 B200_TRACE=1 python tools/probe_content.py 256 grain 2>&1 | grep -A6 "^band" | head -8 >> $L
cat $L
