#!/bin/bash
# usage: r02_ncu.sh tag [kernel regex] [extra bench args]: one `ncu --set full` capture of the fused kernel (4th launch), report under gpurun_out/
tag=$1; k=${2:-amp_trim_pileup_warp}; shift; shift
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --lean --e2e-steps 1 $@"
$B > /dev/null 2>&1 || { echo "bench failed without ncu"; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/r02_${tag} $B > gpurun_out/r02_${tag}_ncu.log 2>&1
tail -3 gpurun_out/r02_${tag}_ncu.log
ls -la gpurun_out/r02_${tag}.ncu-rep
