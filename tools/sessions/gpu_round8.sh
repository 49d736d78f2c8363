#!/bin/bash
# Driver-like verification: the GPU test command in ONE process, smoke, both bench arms with default flags.
mkdir -p gpurun_out/final
echo "#### pytest -m gpu (one process, as the driver runs it)"; SECONDS=0; timeout 2400 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4; echo "seconds: $SECONDS"
echo "#### smoke"; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "#### bench (default flags)"; timeout 900 python bench.py 2>gpurun_out/final/bench_default.err | tail -1 | tee gpurun_out/final/bench_default.json | cut -c1-400; tail -2 gpurun_out/final/bench_default.err
echo "#### bench --impl reference (default flags)"; SECONDS=0; timeout 900 python bench.py --impl reference 2>/dev/null | tail -1 | tee gpurun_out/final/bench_reference.json | cut -c1-400; echo "seconds: $SECONDS"
python - <<'P'
import json
d = json.load(open('gpurun_out/final/bench_default.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'train', d['extra']['train'].get('value'), d['extra']['train'].get('ms_per_step'), 'launches', d['gpu_launches'], 'clocks', d['clocks'])
P
