#!/bin/bash
# two ranks as the driver launches them: b200 arm, then the reference arm
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r4b_n$N.json 2> gpurun_out/r4b_n$N.err
echo "rc=$?"; tail -5 gpurun_out/r4b_n$N.err; cut -c1-1200 gpurun_out/r4b_n$N.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 3 --warmup 3 > gpurun_out/r4b_ref_n$N.json 2> gpurun_out/r4b_ref_n$N.err
echo "rc=$?"; cut -c1-400 gpurun_out/r4b_ref_n$N.json
