#!/bin/bash
# final verification of the round-2 tree: driver-like test run, smoke, both bench arms, config 3
mkdir -p gpurun_out/final
echo "#### pytest -m gpu (one process)"; SECONDS=0; timeout 2400 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -2; echo "seconds: $SECONDS"
echo "#### smoke"; python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -1
echo "#### config 3"; timeout 600 python tools/config_bench.py fusion 2>&1 | tail -1 | tee gpurun_out/final/config3_verify.json
echo "#### bench"; timeout 900 python bench.py 2>/dev/null | tail -1 | tee gpurun_out/final/bench_verify.json | cut -c1-260
echo "#### bench --impl reference"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-260
