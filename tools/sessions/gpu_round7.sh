#!/bin/bash
# GPU session 7 of round 2: native L-BFGS driver (csrc/lbfgs.cu) + end-to-end workflow test.
mkdir -p gpurun_out/final
echo "#### fusion tests"; timeout 900 python -m pytest tests/test_fusion_gpu.py -x -q -m gpu 2>&1 | tail -6
echo "#### e2e workflows"; timeout 900 python -m pytest tests/test_e2e_flows_gpu.py -x -q -m gpu 2>&1 | tail -25
for w in 1 4 8; do
  echo "#### config 3 (UNet half), native, $w workers"; MOS_FUSION_WORKERS=$w timeout 900 python tools/config_bench.py fusion 2>&1 | tail -1 | tee gpurun_out/final/config3_native_w$w.json
done
echo "#### compose_concepts SD1.5 size"; timeout 1200 python tools/compose_bench.py 2>gpurun_out/final/compose.err | tail -1 | tee gpurun_out/final/compose_sd15_native.json | cut -c1-1500; tail -2 gpurun_out/final/compose.err | cut -c1-300
echo "#### ncu of the fusion kernels"; timeout 300 ncu --set full --clock-control none -k regex:'dgemm_mixed|lbfgs_step' -c 6 -o /tmp/fus -f python tools/ncu_targets.py > /dev/null 2>&1; python tools/ncu_summary.py /tmp/fus.ncu-rep gpurun_out/final/r2_fusion_kernels 2>&1 | tail -1
