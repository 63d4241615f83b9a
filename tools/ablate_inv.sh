#!/bin/bash
# ablation timing of the packed c2 inverse kernel (needs a -DVEILS_ABLATE build of the library)
cp veils_b200/build/abl/libveils_cuda_ablate.so veils_b200/libveils_cuda.so
for d in ${ABL_LIST:-0 1 4 8 16 5 12 13 29}; do
  echo "== VEILS_PK_DBG_INV=$d (1: no global loads, 2: no shuffles, 4: no gather, 8: no last pass, 16: no pass-1 arithmetic)"
  VEILS_PK_DBG_INV=$d python tools/time_configs.py --align "c2 f32" 2>&1 | tail -1
done
