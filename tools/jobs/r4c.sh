#!/bin/bash
# NOTE: the POLAR_SC5_POL experiment switch this script drives was removed again after the measurement (commit "OSD: 16-byte aligned ..."); kept as the record of what DESIGN 4.1 quotes
python -m pytest tests/test_gpu_osd.py -m gpu -x -q 2>&1 | tail -15
M=dram__bytes_write.sum,dram__bytes_read.sum,gpu__time_duration.sum
for pol in 0 2 3 12 13 22 200 100 222 322; do
  echo "== POLAR_SC5_POL=$pol"
  POLAR_SC5_POL=$pol ncu --metrics $M --clock-control none -k regex:sc5_kernel -s 1 -c 1 python tools/ncu_target.py sc 2 2>&1 | grep -E "dram__|gpu__time" | tr -s ' ' | tr '\n' ';'; echo
  POLAR_SC5_POL=$pol python tools/sc_check.py 1024 2>&1 | tail -1
done
