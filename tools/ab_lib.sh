#!/bin/bash
# A/B timing of an experimental build of the library against the production build:
#   tools/ab_lib.sh <path/to/experimental.so> [bench args]
exp=$1; shift
echo "== production"; python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e "$@" | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('value', round(d['value']), 'stft', r['stft'], 'istft', r['istft'], 'err', d['recon_max_abs_err'])"
cp veils_b200/libveils_cuda.so /tmp/prod.so; cp $exp veils_b200/libveils_cuda.so
echo "== $exp"; python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e "$@" | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('value', round(d['value']), 'stft', r['stft'], 'istft', r['istft'], 'err', d['recon_max_abs_err'])"
cp /tmp/prod.so veils_b200/libveils_cuda.so
