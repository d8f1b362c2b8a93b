#!/bin/bash
# round-2 experiment C: SM partition (green contexts) variants against the unpartitioned pipeline, B = 128 and 64
mkdir -p gpurun_out
L=gpurun_out/r2c.log
: > $L
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) >> $L
run() { echo "== $*" >> $L; env "$@" B200_VERBOSE=1 python tools/probe_content.py ${B:-128} grain 2>&1 | grep -v "^B=.*grain" | head -3 >> $L; env "$@" python tools/probe_content.py ${B:-128} grain 2>&1 | grep "^B=" | tail -1 >> $L; }
run B200_NO_PARTITION=1
run B200_RANGE_SMS=8
run B200_RANGE_SMS=16
run B200_RANGE_SMS=24
run B200_RANGE_SMS=32
run B200_RANGE_SMS=16 B200_KPAR=3
run B200_RANGE_SMS=8 B200_EMIT_SMS=16
run B200_RANGE_SMS=16 B200_EMIT_SMS=16
run B200_RANGE_SMS=16 B200_EMIT_SMS=24 B200_KPAR=3
run B200_RANGE_SMS=32 B200_EMIT_MODE=2
run B200_RANGE_SMS=40 B200_EMIT_MODE=2 B200_KPAR=3
B=64 run B200_NO_PARTITION=1
B=64 run B200_RANGE_SMS=8
B=64 run B200_RANGE_SMS=16
B=192 run B200_RANGE_SMS=16
B=192 run B200_RANGE_SMS=24
cat $L
