#!/bin/bash
# last pass: GPU suite, counters for the final sources, both bench arms, smoke
python -m pytest tests -m gpu -x -q > gpurun_out/r5k_tests.log 2>&1; tail -2 gpurun_out/r5k_tests.log
python tools/refresh_counters.py > gpurun_out/r5k_counters.log 2>&1; tail -2 gpurun_out/r5k_counters.log
cp gpurun_out/sc_counters.json profiles/sc_counters.json
python bench.py --steps 20 --warmup 5 > gpurun_out/r5k_bench.json 2> gpurun_out/r5k_bench.err; echo bench rc=$?
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r5k_ref.json 2> gpurun_out/r5k_ref.err; echo ref rc=$?
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
ncu --set full --clock-control none --import-source on -k regex:scl3_kernel -s 1 -c 1 -f -o gpurun_out/r02_scl3 python tools/ncu_target.py scl 2 > gpurun_out/r5k_ncu_scl.log 2>&1; echo "ncu scl rc=$?"
