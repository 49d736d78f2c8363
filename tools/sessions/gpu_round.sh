#!/bin/bash
# One GPU session of round 2: (1) GroupNorm one-pass vs two-pass per call inside the engine, (2) the whole GPU suite with
# the two-launch GroupNorm (known-good), (3) the suite files that exercise GroupNorm with the one-pass kernel, (4) benches.
mkdir -p gpurun_out
echo "#### gn_debug"; python tools/gn_debug.py 2>&1 | tail -16
echo "#### suite (two-pass GN)"; MOS_GN_TWOPASS=1 tools/gpu_suite.sh
grep -h "rel-L2\|FAILED\|Error\|error" gpurun_out/suite/test_regional_gpu.log gpurun_out/suite/test_trainer_full_gpu.log gpurun_out/suite/test_vae_gpu.log gpurun_out/suite/test_unet_gpu.log | head -40
mkdir -p gpurun_out/suite_twopass; cp gpurun_out/suite/*.log gpurun_out/suite_twopass/
echo "#### unet / dropin / regional (one-pass GN)"; tools/gpu_suite.sh tests/test_unet_gpu.py tests/test_dropin_gpu.py tests/test_regional_gpu.py
echo "#### bench two-pass"; MOS_GN_TWOPASS=1 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_twopass.json | cut -c1-1200
echo "#### bench one-pass"; python bench.py --no-cpu-baseline --no-train 2>&1 | tail -1 | tee gpurun_out/bench_onepass.json | cut -c1-400
echo "#### ncu launch list of one eager step"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/r2_launches.csv python tools/profile_step.py --runs 3 > gpurun_out/r2_launches.log 2>&1; tail -2 gpurun_out/r2_launches.log
echo "#### ncu --set full of the per-family targets"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'attn_kernel|attn_bwd|gn_group|gn_stats|gn_apply|layernorm|lora_grad|dgemm_mixed|gemm_kernel|splitk|softmax_rows' -o gpurun_out/r2_kernels -f python tools/ncu_targets.py > gpurun_out/r2_kernels.log 2>&1; tail -3 gpurun_out/r2_kernels.log
ls -la gpurun_out/*.ncu-rep 2>/dev/null
echo "#### step breakdown"; timeout 300 python tools/step_breakdown.py 2>&1 | tail -12
