#!/bin/bash
# train_edlora.py -opt <yml> under torchrun on 2 GPUs (16 samples / (2 GPUs x batch 2) = 4 optimiser steps), rank 0 saves
set -o pipefail
python tools/dp_train_e2e.py prepare /tmp/dp_e2e 2>&1 | tail -1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 mix-of-show_b200/train_edlora.py -opt /tmp/dp_e2e/train.yml 2>&1 | grep -v "OMP_NUM_THREADS\|\*\*\*\*\*\|Warning\|warn" | tail -12
python tools/dp_train_e2e.py check /tmp/dp_e2e
echo "#### quick single-GPU sanity of the final tree"; timeout 600 python -m pytest tests/test_dropin_gpu.py tests/test_e2e_flows_gpu.py -x -q -m gpu 2>&1 | tail -1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
