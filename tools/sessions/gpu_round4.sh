#!/bin/bash
# GPU session 4 of round 2: final defaults (suite, bench, launch list), pipeline-depth A/B, compose_concepts end to end, smoke.
mkdir -p gpurun_out/final
b() { # name, extra bench args..., env via BENV
  local name=$1; shift
  env $BENV timeout 600 python bench.py --no-cpu-baseline --no-train "$@" 2>/dev/null | tail -1 > gpurun_out/final/$name.json
  python - "$name" <<'P'
import json, sys
n = sys.argv[1]
try:
    d = json.load(open(f'gpurun_out/final/{n}.json'))
    print(f"{n:28s} ms_per_step {d['ms_per_step']:.3f}  value {d['value']:.1f}  e2e {d['e2e']['value']:.1f}  launches {d['gpu_launches']}  frac {d['roofline']['frac']:.3f}")
except Exception as e:
    print(n, 'FAILED', e)
P
}
echo "#### smoke"; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "#### A/B pipeline depth"
BENV="MOS_GEMM_STAGES=0" b stages_default
BENV="MOS_GEMM_STAGES=3" b stages_3
BENV="MOS_GEMM_STAGES=0" b stages_default2
BENV="MOS_GEMM_STAGES=3" b stages_3b
BENV="MOS_GEMM_STAGES=0" b images4 --images 4
echo "#### compose_concepts end to end (tiny first: fast failure)"; timeout 600 python tools/compose_bench.py --tiny --concepts 2 --textenc-iters 20 --unet-iters 5 2>&1 | tail -1 | cut -c1-1200
timeout 1500 python tools/compose_bench.py 2>gpurun_out/final/compose.err | tail -1 | tee gpurun_out/final/compose_sd15.json | cut -c1-1500; tail -3 gpurun_out/final/compose.err
echo "#### launch list with the final defaults"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1700 --csv --log-file gpurun_out/final/r2_launches.csv python tools/profile_step.py --runs 3 > gpurun_out/final/r2_launches.log 2>&1; tail -1 gpurun_out/final/r2_launches.log
echo "#### step breakdown"; timeout 300 python tools/step_breakdown.py 2>&1 | tail -8 | tee gpurun_out/final/step_breakdown.txt
echo "#### full bench"; timeout 900 python bench.py 2>gpurun_out/final/bench_full.err | tail -1 | tee gpurun_out/final/bench_full.json | cut -c1-600
