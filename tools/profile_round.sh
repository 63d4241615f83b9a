#!/bin/bash
# Evidence run for profiles/ (run on the B200 through gpurun, one GPU):
#   1. plain bench (must exit 0)           -> gpurun_out/${1}_bench.log
#   2. launch list of a short bench        -> gpurun_out/${1}_launches.csv
#   3. ncu --set full of the two hot kernels (one launch each, after warm-up)
#   usage: tools/profile_round.sh <prefix>
p=${1:-rX}
mkdir -p gpurun_out
python bench.py > gpurun_out/${p}_bench.log 2> gpurun_out/${p}_bench.err || { echo "bench failed"; tail -5 gpurun_out/${p}_bench.err; exit 1; }
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${p}_ref.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/${p}_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${p}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/${p}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pk_ -s 8 -c 2 -f -o gpurun_out/${p}_full \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/${p}_ncu2.log 2>&1
python tools/time_configs.py --align > gpurun_out/${p}_sweep.log 2>&1
tail -2 gpurun_out/${p}_ncu2.log; cat gpurun_out/${p}_sweep.log
