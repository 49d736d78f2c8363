#!/bin/bash
echo "#### fusion + e2e tests"; timeout 1200 python -m pytest tests/test_fusion_gpu.py tests/test_e2e_flows_gpu.py -x -q -m gpu 2>&1 | tail -3
echo "#### config 3"; MOS_FUSION_PROFILE=1 timeout 600 python tools/config_bench.py fusion 2>&1 | tail -2 | tee gpurun_out/final/config3_graphrec.txt
echo "#### compose"; timeout 1200 python tools/compose_bench.py 2>/dev/null | tail -1 | tee gpurun_out/final/compose_sd15_final2.json | cut -c1-1200
