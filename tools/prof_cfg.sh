#!/bin/bash
# ncu --set full of one config's forward+inverse kernels at a small batch
#   usage: tools/prof_cfg.sh <prefix> <config> <batch> <kernel-regex>
p=$1; cfg=$2; b=$3; rx=$4
mkdir -p gpurun_out
python tools/prof_run.py --config $cfg --batch $b --ldp-align 16 > gpurun_out/${p}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${p}_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:$rx -s 4 -c 2 -f -o gpurun_out/${p}_prof \
    python tools/prof_run.py --config $cfg --batch $b --ldp-align 16 > gpurun_out/${p}_ncu.log 2>&1
tail -3 gpurun_out/${p}_ncu.log
