#!/bin/bash
python -m pytest tests/test_gpu_link.py -m gpu -x -q 2>&1 | tail -15
python bench.py --steps 5 --warmup 3 --skip-scl --skip-e2e --skip-cpu --skip-link > gpurun_out/r4e_bench.json 2> gpurun_out/r4e_bench.err; echo bench rc=$?; tail -3 gpurun_out/r4e_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r4e_bench.json').read().strip().splitlines()[-1])
for k,v in d['sweep'].items(): print(k, v['n_gpus'], v['wall_ms'], v['codewords_per_s'], v['iterations_counted'], v['iterations_queued'], v['decoder_launches'], v['split_us_per_iteration'])
P
