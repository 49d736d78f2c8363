#!/bin/bash
# Runs every GPU test file in its OWN process (a CUDA fault in one file must not poison the others) and keeps the logs.
# usage: tools/gpu_suite.sh [file ...]   (default: all tests/test_*gpu*.py); logs -> gpurun_out/suite/
mkdir -p gpurun_out/suite
files=("$@")
if [ ${#files[@]} -eq 0 ]; then files=(tests/test_*gpu*.py); fi
for f in "${files[@]}"; do
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -m gpu -x -q -s > gpurun_out/suite/$n.log 2>&1
  echo "== $n rc=$? : $(grep -E 'passed|failed|error' gpurun_out/suite/$n.log | tail -1)"
done
