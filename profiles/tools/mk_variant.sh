#!/bin/bash
# usage: build/mk.sh name [extra nvcc flags]: build/libamp_<name>.so from the working tree
n=$1; shift
cd /root/repo/amplipy_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC -ccbin /usr/bin/g++ "$@" -o /root/repo/build/libamp_$n.so amp_abi.cu && echo built $n
