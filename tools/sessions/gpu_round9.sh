#!/bin/bash
# fp64 closure GEMM tile shapes under 8 concurrent solves
mkdir -p gpurun_out/final
echo "#### dgemm / driver tests"; timeout 600 python -m pytest tests/test_fusion_gpu.py -x -q -m gpu -k "dgemm or native or direction" 2>&1 | tail -2
for t in 0 1 2 3; do
  echo "#### config 3, MOS_DGEMM_TILE=$t"; MOS_DGEMM_TILE=$t timeout 600 python tools/config_bench.py fusion 2>&1 | tail -1 | tee gpurun_out/final/config3_tile$t.json
done
echo "#### vae pil test"; timeout 300 python -m pytest tests/test_vae_gpu.py -x -q -m gpu 2>&1 | tail -2
