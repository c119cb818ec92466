#!/bin/bash
# Round-end validation on the GPU box: smoke(), the GPU test-suite, the timing build's phase split.  usage: r02_validate.sh tag
tag=$1
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_${tag}_smoke.log 2>&1; echo "smoke exit $?"; tail -4 gpurun_out/r02_${tag}_smoke.log
python -m pytest tests -m gpu -x -q > gpurun_out/r02_${tag}_pytest.log 2>&1; tail -3 gpurun_out/r02_${tag}_pytest.log
if [ -f build/libamp_timing.so ]; then
AMP_LIB_OVERRIDE=$PWD/build/libamp_timing.so python bench.py --steps 2 --warmup 3 --lean --e2e-steps 1 > /dev/null 2> gpurun_out/r02_${tag}_timing.err
head -6 gpurun_out/r02_${tag}_timing.err | tail -3
fi
