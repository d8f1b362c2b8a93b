#!/bin/bash
# what slows k_range down inside the pipeline (13.4 ms per band against 8.1 alone)?
mkdir -p gpurun_out
L=gpurun_out/r2y.log
: > $L
tr() { echo "== $*" >> $L; env "$@" B200_TRACE=1 python tools/probe_content.py 128 grain 2>&1 | grep -A5 "^band" | tail -3 >> $L; env "$@" python tools/probe_content.py 128 grain 2>&1 | grep "^B=" | tail -1 >> $L; }
tr X=1
tr B200_SKIP_EMIT=1
tr B200_EMIT_MODE=2
tr B200_RANGE_SMS=32
tr B200_RANGE_SMS=16 B200_RANGE_CTAS_PER_SM=2
tr B200_KPAR=3
cat $L
