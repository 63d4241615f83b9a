#!/bin/bash
# ablation timing of the packed c2 forward kernel (needs a -DVEILS_ABLATE build of the library)
cp veils_b200/build/abl/libveils_cuda_ablate.so veils_b200/libveils_cuda.so
for d in ${ABL_LIST:-0 1 4 8 5 9 12}; do
  echo "== VEILS_PK_DBG=$d (1: no global stores, 2: no shuffles (plain-store kernels), 4: no pass 1, 8: no last pass)"
  VEILS_PK_DBG=$d python tools/time_configs.py --align "c2 f32" 2>&1 | tail -1
done
