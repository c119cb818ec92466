#!/bin/bash
# usage: r02_ont.sh tag [reads...]: ONT-like bench lines (kernel ms / fraction of the HBM peak) for the given read counts
tag=$1; shift
mkdir -p gpurun_out
for n in "${@:-100000 300000}"; do
  python bench.py --workload ont --reads $n --steps 10 --lean > gpurun_out/r02_${tag}_ont_$n.json 2> gpurun_out/r02_${tag}_ont_$n.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r02_${tag}_ont_$n.json').read().strip().splitlines()[-1])
print('${tag} ont', $n, 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'frac', round(d['roofline']['frac'],4), 'step', round(d['ms_per_step'],4), 'err', d['device_error_flags'], 'chk', d['depth_checksum'])
PY
done
