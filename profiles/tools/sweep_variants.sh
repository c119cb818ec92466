#!/bin/bash
# usage: build/sweep.sh name1 name2 ... : runs a short bench with build/libamp_<name>.so in place of the product library
cp amplipy_b200/csrc/libamplipy_b200.so /tmp/orig.so
for n in "$@"; do
  cp build/libamp_$n.so amplipy_b200/csrc/libamplipy_b200.so
  python bench.py --steps 10 --warmup 3 --lean --e2e-steps 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$n', 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'frac', round(d['roofline']['frac'],4), 'err', d['device_error_flags'], 'chk', d['depth_checksum'])"
done
cp /tmp/orig.so amplipy_b200/csrc/libamplipy_b200.so
