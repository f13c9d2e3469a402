#!/bin/bash
# final single-GPU pass of the round: GPU suite, counters, both bench arms, launch list, full ncu captures, smoke
python -m pytest tests -m gpu -x -q > gpurun_out/r5a_tests.log 2>&1; tail -3 gpurun_out/r5a_tests.log
python tools/refresh_counters.py > gpurun_out/r5a_counters.log 2>&1; tail -2 gpurun_out/r5a_counters.log; cp gpurun_out/sc_counters.json profiles/sc_counters.json
python bench.py --steps 20 --warmup 5 > gpurun_out/r5a_bench.json 2> gpurun_out/r5a_bench.err; echo bench rc=$?
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r5a_ref.json 2> gpurun_out/r5a_ref.err; echo ref rc=$?
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 2 --warmup 3 --skip-cpu > gpurun_out/r5a_ncu_bench.log 2>&1; echo ncu-list rc=$?
for t in sc:sc5_kernel:r02_sc5_n1024 sc4096:sc5_kernel:r02_sc5_n4096 sc2048:sc5_kernel:r02_sc5_n2048 sc512:sc4_kernel:r02_sc4_n512; do
  IFS=: read what kern out <<< "$t"
  ncu --set full --clock-control none --import-source on -k regex:$kern -s 1 -c 1 -f -o gpurun_out/$out python tools/ncu_target.py $what 2 > gpurun_out/r5a_ncu_$what.log 2>&1; echo "ncu $what rc=$?"
done
python tools/sweep_c5.py gpurun_out/r02_c5_sweep_1gpu.md 2>&1 | grep "n=" | tail -6
