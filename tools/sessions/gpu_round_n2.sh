#!/bin/bash
# 2-GPU session: the driver's N=2 command, the config-5 equivalence check, the reference arm under torchrun.
mkdir -p gpurun_out/n2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "#### bench --gpus 2"; timeout 900 $TR --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 2>gpurun_out/n2/bench.err | tail -1 | tee gpurun_out/n2/bench_n2.json | cut -c1-2500
tail -3 gpurun_out/n2/bench.err
echo "#### dp_check (tiny topology)"; timeout 600 $TR --master-port 29512 tools/dp_check.py 2>gpurun_out/n2/dp.err | tail -1 | tee gpurun_out/n2/dp_check_tiny.json; tail -3 gpurun_out/n2/dp.err
echo "#### dp_check (SD1.5 topology)"; timeout 900 $TR --master-port 29513 tools/dp_check.py --full 2>gpurun_out/n2/dp_full.err | tail -1 | tee gpurun_out/n2/dp_check_sd15.json; tail -3 gpurun_out/n2/dp_full.err
echo "#### reference arm under torchrun"; timeout 600 $TR --master-port 29514 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 2>/dev/null | tail -1 | cut -c1-600
