#!/bin/bash
# GPU session 2 of round 2: A/B of the in-kernel split-K finalize and of the L2 weight prefetch, GEMM tests, configs 3 / 4.
mkdir -p gpurun_out/ab
echo "#### gemm tests"; timeout 900 python -m pytest tests/test_gemm_gpu.py -x -q -m gpu 2>&1 | tail -4
b() { # name, env...
  local name=$1; shift
  env "$@" timeout 600 python bench.py --no-cpu-baseline --no-train 2>/dev/null | tail -1 > gpurun_out/ab/$name.json
  python - "$name" <<'P'
import json, sys
n = sys.argv[1]
try:
    d = json.load(open(f'gpurun_out/ab/{n}.json'))
    print(f"{n:28s} ms_per_step {d['ms_per_step']:.3f}  e2e {d['e2e']['value']:.1f}  launches {d['gpu_launches']}  frac {d['roofline']['frac']:.3f}")
except Exception as e:
    print(n, 'FAILED', e)
P
}
echo "#### A/B benches"
b base_fused_nopf MOS_SPLITK_FUSED=1 MOS_L2_PREFETCH=0
b unfused_nopf    MOS_SPLITK_FUSED=0 MOS_L2_PREFETCH=0
b fused_pf        MOS_SPLITK_FUSED=1 MOS_L2_PREFETCH=1
b unfused_pf      MOS_SPLITK_FUSED=0 MOS_L2_PREFETCH=1
b base_again      MOS_SPLITK_FUSED=1 MOS_L2_PREFETCH=0
echo "#### unet parity with prefetch on"; MOS_L2_PREFETCH=1 timeout 900 python -m pytest tests/test_unet_gpu.py tests/test_dropin_gpu.py -x -q -m gpu 2>&1 | tail -3
echo "#### config 4"; timeout 900 python tools/config_bench.py regional 2>&1 | tail -1 | tee gpurun_out/ab/config4.json
echo "#### config 3"; timeout 1200 python tools/config_bench.py fusion 2>&1 | tail -1 | tee gpurun_out/ab/config3.json
echo "#### ncu launch list of one eager step"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/r2_launches.csv python tools/profile_step.py --runs 3 > gpurun_out/r2_launches.log 2>&1; tail -2 gpurun_out/r2_launches.log
echo "#### ncu --set full of the per-family targets (report stays on the box: > 64 MiB; summaries come back)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'attn_kernel|attn_bwd|gn_group|gn_stats|gn_apply|layernorm|lora_grad|dgemm_mixed|gemm_kernel|splitk|softmax_rows' -o /tmp/r2_kernels -f python tools/ncu_targets.py > gpurun_out/r2_kernels.log 2>&1; tail -2 gpurun_out/r2_kernels.log
python tools/ncu_summary.py /tmp/r2_kernels.ncu-rep gpurun_out/r2_kernels 2>&1 | tail -3
echo "#### small report (GEMM conv, both tile modes, + d=40 attention) for the source view"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'attn_kernel|gemm_kernel' -c 8 -o gpurun_out/r2_top -f python tools/ncu_targets.py > /dev/null 2>&1; ls -la gpurun_out/*.ncu-rep
du -sh gpurun_out
echo "#### gemm stage sweep"; timeout 300 python tools/gemm_stage_sweep.py 2>&1 | tail -16
