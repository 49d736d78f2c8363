#!/bin/bash
mkdir -p gpurun_out/final
timeout 600 python bench.py --act-dtype bf16 --no-cpu-baseline --no-train 2>/dev/null | tail -1 | tee gpurun_out/final/bench_bf16_final.json | cut -c1-400
