#!/bin/bash
# ncu --set full of a prof_run.py invocation:  tools/prof_any.sh <prefix> <kernel-regex> <prof_run args...>
p=$1; rx=$2; shift 2
mkdir -p gpurun_out
python tools/prof_run.py "$@" > gpurun_out/${p}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${p}_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:$rx -s 4 -c 2 -f -o gpurun_out/${p}_prof \
    python tools/prof_run.py "$@" > gpurun_out/${p}_ncu.log 2>&1
tail -2 gpurun_out/${p}_ncu.log
