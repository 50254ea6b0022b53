#!/bin/bash
# per-kernel times of the default workload (no e2e / extras)
python bench.py --steps 20 --warmup 5 --no-extra-configs --no-cpu-baseline --no-fp32 --no-e2e 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('step ms', round(d['ms_per_step'],4), 'flow-only ms', round(d['spectral_step']['ms_per_step'],4))
print({k:round(v['ms_avg'],4) for k,v in d['kernels'].items()})"
