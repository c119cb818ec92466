#!/bin/bash
# usage: r02_ncu_any.sh tag kernel_regex skip -- command...: one `ncu --set full` capture of a kernel of any command
tag=$1; k=$2; skip=$3; shift; shift; shift; shift
mkdir -p gpurun_out
"$@" > /dev/null 2>&1 || { echo "command failed without ncu"; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o gpurun_out/r02_${tag} "$@" > gpurun_out/r02_${tag}_ncu.log 2>&1
tail -2 gpurun_out/r02_${tag}_ncu.log
