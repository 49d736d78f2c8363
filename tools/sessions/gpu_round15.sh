#!/bin/bash
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
b epi8 MOS_GEMM_EPI_WARPS=8
b epi4 MOS_GEMM_EPI_WARPS=4
b epi8b MOS_GEMM_EPI_WARPS=8
