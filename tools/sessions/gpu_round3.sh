#!/bin/bash
# GPU session 3 of round 2: epilogue staging tile overlaid on the pipeline stages (6 instead of 4 stages): sweep + A/B + parity.
mkdir -p gpurun_out/ab3
echo "#### stage sweep, staging behind the stages (MOS_GEMM_STG_ALIAS=0)"; MOS_GEMM_STG_ALIAS=0 timeout 300 python tools/gemm_stage_sweep.py 2>&1 | tail -17
echo "#### stage sweep, staging overlays the stages"; MOS_GEMM_STG_ALIAS=1 timeout 300 python tools/gemm_stage_sweep.py 2>&1 | tail -17
echo "#### gemm / kernel tests (alias on)"; timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_f16_kernels_gpu.py -x -q -m gpu 2>&1 | tail -3
b() { # name, env...
  local name=$1; shift
  env "$@" timeout 600 python bench.py --no-cpu-baseline --no-train 2>/dev/null | tail -1 > gpurun_out/ab3/$name.json
  python - "$name" <<'P'
import json, sys
n = sys.argv[1]
try:
    d = json.load(open(f'gpurun_out/ab3/{n}.json'))
    print(f"{n:28s} ms_per_step {d['ms_per_step']:.3f}  e2e {d['e2e']['value']:.1f}  launches {d['gpu_launches']}  frac {d['roofline']['frac']:.3f}")
except Exception as e:
    print(n, 'FAILED', e)
P
}
echo "#### A/B benches"
b alias_off MOS_GEMM_STG_ALIAS=0
b alias_on  MOS_GEMM_STG_ALIAS=1
b alias_off2 MOS_GEMM_STG_ALIAS=0
b alias_on2  MOS_GEMM_STG_ALIAS=1
echo "#### full suite (defaults)"; tools/gpu_suite.sh
grep -h "rel-L2" gpurun_out/suite/test_unet_gpu.log gpurun_out/suite/test_regional_gpu.log | head
echo "#### full bench (train leg = full ED-LoRA step)"; timeout 900 python bench.py 2>gpurun_out/bench_full.err | tail -1 | tee gpurun_out/bench_full.json | cut -c1-3000
tail -5 gpurun_out/bench_full.err
