#!/bin/bash
# usage: build/ont_ab.sh name...: ONT bench with build/libamp_<name>.so in place of the product library
cp amplipy_b200/csrc/libamplipy_b200.so /tmp/orig.so
for n in "$@"; do
  cp build/libamp_$n.so amplipy_b200/csrc/libamplipy_b200.so
  timeout 200 python bench.py --workload ont --reads 300000 --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$n', 'step', round(d['ms_per_step'],3), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'err', d['device_error_flags'], 'chk', d['depth_checksum'])"
done
cp /tmp/orig.so amplipy_b200/csrc/libamplipy_b200.so
