set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/time_configs.py --align > gpurun_out/r2r_sweep.log 2>&1; cat gpurun_out/r2r_sweep.log
tools/prof_any.sh r2r_mr400 'smooth_' --config c2 --m 400 --hop 160 --batch 256 --ldp-align 16
tools/prof_any.sh r2r_f64c2 'pow2_' --config c2 --prec f64 --batch 128 --ldp-align 8
python bench.py > gpurun_out/r2r_bench_1gpu.json 2> gpurun_out/r2r_bench_1gpu.err; tail -2 gpurun_out/r2r_bench_1gpu.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2r_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2r_ncu1.log 2>&1
tail -2 gpurun_out/r2r_ncu1.log | cut -c1-300
