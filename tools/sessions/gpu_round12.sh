#!/bin/bash
# split-K granularity A/B + the final bench lines of the round
mkdir -p gpurun_out/final
b() { local name=$1; shift; env "$@" timeout 600 python bench.py --no-cpu-baseline --no-train 2>/dev/null | tail -1 > gpurun_out/final/$name.json
  python - "$name" <<'P'
import json, sys
n = sys.argv[1]
try:
    d = json.load(open(f'gpurun_out/final/{n}.json'))
    print(f"{n:28s} ms_per_step {d['ms_per_step']:.3f}  value {d['value']:.1f}  e2e {d['e2e']['value']:.1f}  launches {d['gpu_launches']}")
except Exception as e:
    print(n, 'FAILED', e)
P
}
echo "#### split-K granularity"
b sk8 MOS_SPLITK_MIN_KB=8
b sk4 MOS_SPLITK_MIN_KB=4
b sk12 MOS_SPLITK_MIN_KB=12
b sk16 MOS_SPLITK_MIN_KB=16
b sk8b MOS_SPLITK_MIN_KB=8
echo "#### unet parity for the variants"; MOS_SPLITK_MIN_KB=4 timeout 600 python -m pytest tests/test_unet_gpu.py -x -q -m gpu 2>&1 | tail -1
echo "#### full bench (defaults)"; timeout 900 python bench.py 2>gpurun_out/final/bench_full.err | tail -1 | tee gpurun_out/final/bench_full_final.json | cut -c1-300
echo "#### step breakdown + launch list"; timeout 300 python tools/step_breakdown.py 2>&1 | tail -8 | tee gpurun_out/final/step_breakdown.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1700 --csv --log-file gpurun_out/final/r2_launches.csv python tools/profile_step.py --runs 3 > gpurun_out/final/r2_launches.log 2>&1; tail -1 gpurun_out/final/r2_launches.log
