set -x
tools/prof_any.sh r2i_mr400 'smooth_' --config c2 --m 400 --hop 160 --batch 256 --ldp-align 16
tools/prof_any.sh r2i_f64c2 'pow2_' --config c2 --prec f64 --batch 128 --ldp-align 8
tools/prof_any.sh r2i_c5 'pow2_' --config c5 --batch 8 --spec
tools/prof_any.sh r2i_c1 'pow2_' --config c1 --batch 1
python tools/time_configs.py --align > gpurun_out/r2i_sweep.log 2>&1
cat gpurun_out/r2i_sweep.log
