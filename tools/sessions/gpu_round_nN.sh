#!/bin/bash
# N-GPU session: the driver's scaling command at N = $1
N=${1:-4}
mkdir -p gpurun_out/n$N
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "#### bench --gpus $N"; timeout 900 $TR --master-port 29521 bench.py --gpus $N --steps 20 --warmup 3 2>gpurun_out/n$N/bench.err | tail -1 | tee gpurun_out/n$N/bench_n$N.json | cut -c1-300
tail -3 gpurun_out/n$N/bench.err
python - $N <<'P'
import json, sys
n = sys.argv[1]
d = json.load(open(f'gpurun_out/n{n}/bench_n{n}.json'))
t = d['extra']['train']
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], '| train', t.get('value'), t.get('ms_per_step'), 'allreduce_us', t.get('allreduce_us'), 'identical', t.get('params_bit_identical_across_ranks'))
P
