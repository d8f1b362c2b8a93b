#!/bin/bash
# debug: run the config-1 drop-in scenario on the GPU box and bring the artefacts back
set -x
T=/tmp/dbgcli; rm -rf $T; mkdir -p $T; cd $T
python - <<'PY'
import sys,os
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from rawcooked_b200 import synth as S
os.makedirs('seq')
for i in range(10):
    p=S.synth_payload(640,480,S.DPX_RGB_8,1000+i,"grain")
    open('seq/f_%06d.dpx'%i,'wb').write(S.dpx_file(640,480,S.DPX_RGB_8,p,i))
PY
ulimit -c 0
/root/repo/oracle/_ref/rawcooked --all -y -b /root/repo/rawcooked_b200/b200enc seq > log1.txt 2>&1; echo "rc=$?" >> log1.txt
/root/repo/oracle/_ref/rawcooked --check seq.mkv -o ./ > log2.txt 2>&1; echo "rc=$?" >> log2.txt
/root/repo/oracle/_ref/rawcooked -threads 1 --check seq.mkv -o ./ > log3.txt 2>&1; echo "rc=$?" >> log3.txt
nproc > nproc.txt
mkdir -p /root/repo/gpurun_out/dbgcli
cp -r seq.mkv log*.txt nproc.txt /root/repo/gpurun_out/dbgcli/
tail -5 log1.txt log2.txt log3.txt
