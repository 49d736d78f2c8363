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
