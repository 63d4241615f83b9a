set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/time_configs.py --align > gpurun_out/r2t_sweep.log 2>&1; cat gpurun_out/r2t_sweep.log
tools/prof_any.sh r2t_f64c2 'pow2_' --config c2 --prec f64 --batch 128 --ldp-align 8
tools/prof_any.sh r2t_mr400 'smooth_' --config c2 --m 400 --hop 160 --batch 256 --ldp-align 16
python bench.py > gpurun_out/r2t_bench_1gpu.json 2> gpurun_out/r2t_bench_1gpu.err; tail -2 gpurun_out/r2t_bench_1gpu.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2t_ref.json 2>&1
