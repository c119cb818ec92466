#!/bin/bash
# last validation + captures of the round (GPU box, repository root); every step bounded by `timeout`
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_last_smoke.log 2>&1; echo "smoke exit $?"
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/r02_last_pytest.log 2>&1; tail -2 gpurun_out/r02_last_pytest.log
timeout 400 python bench.py > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err; python profiles/tools/show_bench.py gpurun_out/r02_final_bench.json | tail -9
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_final_reference.json 2>/dev/null; tail -c 300 gpurun_out/r02_final_reference.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_write_launches.csv python profiles/tools/prof_write.py > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:amp_bgzf_deflate_kernel -s 1 -c 1 -f -o gpurun_out/r02_deflate python profiles/tools/prof_write.py > /dev/null 2>&1
ncu -i gpurun_out/r02_deflate.ncu-rep --page raw --csv > gpurun_out/r02_deflate_raw.csv 2>/dev/null
ls -la gpurun_out/r02_deflate_raw.csv gpurun_out/r02_write_launches.csv
