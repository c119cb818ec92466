#!/bin/bash
# usage: r02_ncu_lite.sh name...: instruction count / IPC / cycles of the fused kernel with build/libamp_<name>.so
mkdir -p gpurun_out
for n in "$@"; do
  AMP_LIB_OVERRIDE=$PWD/build/libamp_$n.so ncu --metrics smsp__inst_executed.sum,sm__cycles_active.avg,sm__inst_executed.avg.per_cycle_active,gpu__time_duration.sum,smsp__warps_active.avg.per_cycle_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:amp_trim_pileup_warp -s 3 -c 1 --csv python bench.py --steps 2 --warmup 3 --lean --e2e-steps 1 2>/dev/null | grep -v "^==\|^{" | awk -F, -v n=$n 'NR>1{gsub(/"/,""); print n, $(NF-2), $NF}'
done
