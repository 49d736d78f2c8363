#!/bin/bash
# GPU session 6 of round 2: concurrent per-layer solves (host threads x CUDA streams) + faster fp64 closure GEMM.
mkdir -p gpurun_out/final
echo "#### fusion tests"; timeout 1200 python -m pytest tests/test_fusion_gpu.py -x -q -m gpu 2>&1 | tail -3
echo "#### config 3 (UNet half), 1 worker"; MOS_FUSION_WORKERS=1 timeout 1200 python tools/config_bench.py fusion 2>&1 | tail -1 | tee gpurun_out/final/config3_w1.json
echo "#### config 3 (UNet half), 4 workers"; MOS_FUSION_WORKERS=4 timeout 1200 python tools/config_bench.py fusion 2>&1 | tail -1 | tee gpurun_out/final/config3_w4.json
echo "#### config 3 (UNet half), 8 workers"; MOS_FUSION_WORKERS=8 timeout 1200 python tools/config_bench.py fusion 2>&1 | tail -1 | tee gpurun_out/final/config3_w8.json
echo "#### compose_concepts SD1.5 size (default workers)"; timeout 1800 python tools/compose_bench.py 2>gpurun_out/final/compose.err | tail -1 | tee gpurun_out/final/compose_sd15.json | cut -c1-1500; tail -2 gpurun_out/final/compose.err | cut -c1-300
echo "#### ncu of the fusion kernels"; timeout 300 ncu --set full --clock-control none -k regex:'dgemm_mixed|lbfgs_step' -c 6 -o /tmp/fus -f python tools/ncu_targets.py > /dev/null 2>&1; python tools/ncu_summary.py /tmp/fus.ncu-rep gpurun_out/final/r2_fusion_kernels 2>&1 | tail -1
