#!/bin/bash
# where do the DRAM writes of sc5 (n=1024) come from?  warps per SM x discard on/off, ncu counters + timing
M=dram__bytes_write.sum,dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sector_hit_rate.pct
for cfg in "8 1" "6 1" "4 1" "8 0"; do
  set -- $cfg
  echo "== warps=$1 discard=$2"
  POLAR_SC_WARPS_SM=$1 POLAR_SC4_DISCARD=$2 ncu --metrics $M --clock-control none -k regex:sc5_kernel -s 1 -c 1 python tools/ncu_target.py sc 2 2>&1 | grep -E "dram__|gpu__time|lts__"
done
POLAR_SC3_DBG=1 python tools/sc_check.py 1024 2>&1 | tail -2
