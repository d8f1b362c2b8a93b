#!/bin/bash
# SM split between k_range and k_model now that k_range keeps its speed beside k_model
mkdir -p gpurun_out
L=gpurun_out/r3b.log
: > $L
run() { echo "== B=${B:-128} ${K:-grain} $*" >> $L; env "$@" python tools/probe_content.py ${B:-128} ${K:-grain} 2>&1 | grep "^B=" | tail -1 >> $L; }
run B200_RANGE_SMS=12 B200_RANGE_CTAS_PER_SM=2
run B200_RANGE_SMS=16 B200_RANGE_CTAS_PER_SM=2
run B200_RANGE_SMS=8 B200_RANGE_CTAS_PER_SM=3
run B200_RANGE_SMS=20
B=96 run B200_RANGE_SMS=12 B200_RANGE_CTAS_PER_SM=2
B=96 run B200_RANGE_SMS=18
B=192 run B200_RANGE_SMS=18 B200_RANGE_CTAS_PER_SM=2
B=192 run B200_RANGE_SMS=12 B200_RANGE_CTAS_PER_SM=4
B=256 run B200_RANGE_SMS=16 B200_RANGE_CTAS_PER_SM=4
echo "== trace 12x2" >> $L
B200_RANGE_SMS=12 B200_RANGE_CTAS_PER_SM=2 B200_TRACE=1 python tools/probe_content.py 128 grain 2>&1 | grep -A6 "^band" | tail -3 >> $L
cat $L
