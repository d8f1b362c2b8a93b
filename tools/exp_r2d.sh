#!/bin/bash
# round-2 experiment D: k_range CTAs of 4/8 warps inside the partition; new bench.py; new tests
mkdir -p gpurun_out
L=gpurun_out/r2d.log
: > $L
run() { echo "== B=${B:-128} $*" >> $L; env "$@" python tools/probe_content.py ${B:-128} grain 2>&1 | grep "^B=" | tail -1 >> $L; }
run B200_RANGE_SMS=24
run B200_RANGE_SMS=16
run B200_RANGE_SMS=32
run B200_RANGE_SMS=24 B200_KPAR=3
run B200_RANGE_SMS=16 B200_KPAR=3
run B200_RANGE_SMS=24 B200_EMIT_SMS=16
B=64 run B200_RANGE_SMS=24
B=64 run B200_RANGE_SMS=16
B=192 run B200_RANGE_SMS=24
B=192 run B200_RANGE_SMS=32
echo "== trace R=24" >> $L
B200_TRACE=1 python tools/probe_content.py 128 grain 2>&1 | tail -45 >> $L
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) >> $L
timeout 600 python bench.py > gpurun_out/r2d_bench3.json 2>> $L
timeout 600 python bench.py --config 2 --no-cpu > gpurun_out/r2d_bench2.json 2>> $L
timeout 600 python bench.py --config 5 --no-cpu > gpurun_out/r2d_bench5.json 2>> $L
timeout 600 python bench.py --config 4 --no-cpu > gpurun_out/r2d_bench4.json 2>> $L
cat $L
for f in gpurun_out/r2d_bench*.json; do echo $f; cut -c1-1500 $f; done
