#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_ffv1_gpu.py -m gpu -x -q -k "prefetch or two_host or device_resident or more_items or golden" 2>&1 | tail -5) > gpurun_out/r4h.log
timeout 400 python bench.py --no-decode --cpu-seconds 4 > gpurun_out/r4h_bench.json 2> gpurun_out/r4h_bench.err
tail -2 gpurun_out/r4h_bench.err >> gpurun_out/r4h.log
python -c "
import json; d=json.load(open('gpurun_out/r4h_bench.json')); print('value', d['value'], 'e2e', d['e2e']['value'], d['check'], d['check_detail']['device_batch_equals_host_batch'])" >> gpurun_out/r4h.log
cat gpurun_out/r4h.log
