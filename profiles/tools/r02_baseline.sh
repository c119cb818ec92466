#!/bin/bash
# round-2 baseline: phase timing build, plain bench (config 2 + ONT), run on the GPU box from the repo root
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02_base_smi.txt 2>&1
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 3 > gpurun_out/r02_base_bench.json 2> gpurun_out/r02_base_bench.err
AMP_LIB_OVERRIDE=$PWD/build/libamp_timing.so python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r02_base_timing.json 2> gpurun_out/r02_base_timing.err
tail -8 gpurun_out/r02_base_timing.err
cat gpurun_out/r02_base_bench.json | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['roofline']['kernel_ms'], d['roofline']['frac'], d['ms_per_step'], d['e2e']['ms_per_step'])"
