#!/bin/bash
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r4f_bench$N.json 2> gpurun_out/r4f_bench$N.err; echo "bench rc=$?"
python - <<P
import json
d=json.loads(open('gpurun_out/r4f_bench$N.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'])
for k,v in d['sweep'].items(): print(k, v['n_gpus'], v['wall_ms'], v['codewords_per_s'], v['iterations_counted'], v['iterations_queued'], v['decoder_launches'], v['split_us_per_iteration'])
P
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 tools/sweep_c5.py gpurun_out/r02_c5_sweep_${N}gpu.md 2>&1 | grep "n=" | tail -6
