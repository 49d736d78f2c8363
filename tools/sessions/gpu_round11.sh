#!/bin/bash
# GroupNorm cluster width A/B + driver-like verification of the final tree
mkdir -p gpurun_out/final
b() { local name=$1; shift; env "$@" timeout 600 python bench.py --no-cpu-baseline --no-train 2>/dev/null | tail -1 > gpurun_out/final/$name.json
  python - "$name" <<'P'
import json, sys
n = sys.argv[1]
try:
    d = json.load(open(f'gpurun_out/final/{n}.json'))
    print(f"{n:28s} ms_per_step {d['ms_per_step']:.3f}  value {d['value']:.1f}  e2e {d['e2e']['value']:.1f}")
except Exception as e:
    print(n, 'FAILED', e)
P
}
echo "#### GroupNorm cluster width"
b gn_148 MOS_GN_MIN_CTAS=148
b gn_296 MOS_GN_MIN_CTAS=296
b gn_592 MOS_GN_MIN_CTAS=592
b gn_148b MOS_GN_MIN_CTAS=148
echo "#### pytest -m gpu (one process)"; SECONDS=0; timeout 2400 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3; echo "seconds: $SECONDS"
echo "#### smoke"; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
