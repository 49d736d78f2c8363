#!/bin/bash
# ring + CUDA-graph L-BFGS direction, 64x64 closure tiles, one engine for all concepts in the spatial stage
mkdir -p gpurun_out/final
echo "#### fusion + e2e tests"; timeout 1200 python -m pytest tests/test_fusion_gpu.py tests/test_e2e_flows_gpu.py -x -q -m gpu 2>&1 | tail -3
echo "#### config 3 (UNet half)"; timeout 600 python tools/config_bench.py fusion 2>&1 | tail -1 | tee gpurun_out/final/config3_final.json
echo "#### config 3 again"; timeout 600 python tools/config_bench.py fusion 2>&1 | tail -1 | tee gpurun_out/final/config3_final2.json
echo "#### compose_concepts SD1.5 size"; timeout 1200 python tools/compose_bench.py 2>gpurun_out/final/compose.err | tail -1 | tee gpurun_out/final/compose_sd15_final.json | cut -c1-1500; tail -2 gpurun_out/final/compose.err | cut -c1-300
