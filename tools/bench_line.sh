#!/bin/bash
# one compact line from bench.py (kernel-only): tools/bench_line.sh [bench args]
python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e "$@" | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('value', round(d['value']), 'stft %.4f ms %.3f' % (r['stft']['ms'], r['stft']['frac']), 'istft %.4f ms %.3f' % (r['istft']['ms'], r['istft']['frac']), 'dense', r.get('dense_pitch'), 'err', d['recon_max_abs_err'])"
