#!/bin/bash
run() { python bench.py --workload ont --reads 300000 --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', 'kernel_ms', round(d['roofline']['kernel_ms'],3), 'frac', round(d['roofline']['frac'],4), 'err', d['device_error_flags'], 'chk', d['depth_checksum'])"; }
run default
AMP_TILE="1024,512,122880,61440,256" run stage_all_256
AMP_TILE="1024,512,61440,30720,128" run stage_all_128
AMP_TILE="1024,1024,30720,15360,64" run stage_all_64
