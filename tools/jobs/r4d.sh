#!/bin/bash
python -m pytest tests/test_gpu_osd.py -m gpu -x -q 2>&1 | tail -5
python tools/batch_scaling.py 2>&1 | tail -9
python tools/refresh_counters.py > gpurun_out/r4d_counters.log 2>&1; tail -3 gpurun_out/r4d_counters.log
