#!/bin/bash
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > gpurun_out/r4d_tests.log
timeout 500 python bench.py --config 2 --cpu-seconds 5 > gpurun_out/r4d_bench2.json 2> gpurun_out/r4d_bench.err
timeout 500 python bench.py > gpurun_out/r4d_bench3.json 2>> gpurun_out/r4d_bench.err
cat gpurun_out/r4d_tests.log; tail -3 gpurun_out/r4d_bench.err; python -c "
import json
for c in (2,3):
    d=json.load(open('gpurun_out/r4d_bench%d.json'%c)); print(c, d['value'], d['e2e']['value'], d['check']); print(json.dumps(d.get('analysis'), indent=1)); print(json.dumps(d.get('decode_check'))[:400])"
