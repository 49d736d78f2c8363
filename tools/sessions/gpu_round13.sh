#!/bin/bash
# last session: refresh the per-kernel ncu summary with the final defaults, driver-like test run, smoke
mkdir -p gpurun_out/final
echo "#### ncu --set full of the per-family targets"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'attn_kernel|attn_bwd|gn_group|gn_stats|gn_apply|layernorm|lora_grad|dgemm_mixed|lbfgs_step|gemm_kernel|splitk|softmax_rows' -c 60 -o /tmp/r2_kernels -f python tools/ncu_targets.py > gpurun_out/final/r2_kernels.log 2>&1; tail -2 gpurun_out/final/r2_kernels.log
python tools/ncu_summary.py /tmp/r2_kernels.ncu-rep gpurun_out/final/r2_kernels 2>&1 | tail -1
echo "#### pytest -m gpu (one process)"; SECONDS=0; timeout 2400 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -2; echo "seconds: $SECONDS"
echo "#### smoke"; python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -1
