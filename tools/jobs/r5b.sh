#!/bin/bash
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -k "scl" 2>&1 | tail -3
python -m pytest tests/test_gpu_link.py -m gpu -x -q -k "mc_control_group or monte_carlo" 2>&1 | tail -2
L=8 SS=54 python tools/perf_probe.py scl3 1024 262144 2>&1 | grep scl3 | tail -1
L=32 SS=54 python tools/perf_probe.py scl3 2048 9472 2>&1 | grep scl3 | tail -1
ncu --set full --clock-control none --import-source on -k regex:scl3_kernel -s 1 -c 1 -f -o gpurun_out/r02_scl3_L32 python tools/ncu_target.py scl32 2 > gpurun_out/r5b_ncu.log 2>&1; echo ncu rc=$?
