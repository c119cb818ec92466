#!/bin/bash
# usage: r02_ont_sweep.sh reads name...: ONT-like bench with build/libamp_<name>.so
n=$1; shift
for v in "$@"; do
  AMP_LIB_OVERRIDE=$PWD/build/libamp_$v.so python bench.py --workload ont --reads $n --steps 10 --lean 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'frac', round(d['roofline']['frac'],4), 'err', d['device_error_flags'], 'chk', d['depth_checksum'])"
done
