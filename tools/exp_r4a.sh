#!/bin/bash
# GPU tests (all), bench line with the decode leg
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/r4a_tests.log
timeout 500 python bench.py > gpurun_out/r4a_bench.json 2> gpurun_out/r4a_bench.err
cat gpurun_out/r4a_tests.log; tail -3 gpurun_out/r4a_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r4a_bench.json')); print(d['value'], d['e2e']['value'], d['check']); print(json.dumps(d.get('decode_check'), indent=1))"
