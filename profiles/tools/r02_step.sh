#!/bin/bash
# usage: r02_step.sh tag [pytest]: bench (+ timing build if build/libamp_timing.so exists) on the GPU box, outputs under gpurun_out/r02_<tag>_*
tag=$1
mkdir -p gpurun_out
if [ "$2" = "pytest" ]; then python -m pytest tests -m gpu -x -q > gpurun_out/r02_${tag}_pytest.log 2>&1; tail -3 gpurun_out/r02_${tag}_pytest.log; fi
python bench.py --steps 20 --warmup 3 --lean --e2e-steps 3 > gpurun_out/r02_${tag}_bench.json 2> gpurun_out/r02_${tag}_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_${tag}_bench.json').read().strip().splitlines()[-1])
print('${tag}', 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'frac', round(d['roofline']['frac'],4), 'step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],3), 'err', d['device_error_flags'], 'chk', d['depth_checksum'])
PY
if [ -f build/libamp_timing.so ]; then
AMP_LIB_OVERRIDE=$PWD/build/libamp_timing.so python bench.py --steps 2 --warmup 3 --lean --e2e-steps 1 > /dev/null 2> gpurun_out/r02_${tag}_timing.err
head -6 gpurun_out/r02_${tag}_timing.err | tail -3
fi
