#!/bin/bash
# quick check after a kernel change: the tests that run the whole model + default bench line with the per-kernel table
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "${1:-whole_model or u8 or eval or reproducible or any_resolution_tiles}" 2>&1 | tail -5
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print('ms_per_step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'parity', d.get('parity_max_abs'))
for k,v in d['kernels'].items(): print(' ', k, v['launches'], v['ms'])
PY
