#!/bin/bash
# GPU session 5 of round 2: fused L-BFGS direction + pipelined fp64 closure GEMM: parity, config 3 timing, compose_concepts.
mkdir -p gpurun_out/final
echo "#### fusion / clip tests"; timeout 1200 python -m pytest tests/test_fusion_gpu.py tests/test_clip_gpu.py -x -q -m gpu 2>&1 | tail -5
echo "#### config 3 (UNet half)"; timeout 1200 python tools/config_bench.py fusion 2>&1 | tail -1 | tee gpurun_out/final/config3_v2.json
echo "#### compose_concepts tiny"; timeout 600 python tools/compose_bench.py --tiny --concepts 2 --textenc-iters 20 --unet-iters 5 2>gpurun_out/final/compose_tiny.err | tail -1 | cut -c1-1200; tail -4 gpurun_out/final/compose_tiny.err | cut -c1-300
echo "#### compose_concepts SD1.5 size"; timeout 1800 python tools/compose_bench.py 2>gpurun_out/final/compose.err | tail -1 | tee gpurun_out/final/compose_sd15.json | cut -c1-1500; tail -4 gpurun_out/final/compose.err | cut -c1-300
echo "#### ncu of the fusion kernels"; timeout 300 ncu --set full --clock-control none -k regex:'dgemm_mixed|lbfgs_step' -c 8 -o /tmp/fus -f python tools/ncu_targets.py > /dev/null 2>&1; python tools/ncu_summary.py /tmp/fus.ncu-rep gpurun_out/final/r2_fusion_kernels 2>&1 | tail -1
