#!/bin/bash
# Final round-2 captures (run on the GPU box from the repository root): launch list of the bench command, one `ncu --set full`
# capture each of the three dominant kernels, exported as CSV / text next to the reports under gpurun_out/.
mkdir -p gpurun_out
B="python bench.py --lean --steps 2 --warmup 3 --e2e-steps 1"
$B > gpurun_out/r02_final_lean.json 2>/dev/null || { echo "bench failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv $B > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:amp_trim_pileup_warp -s 3 -c 1 -f -o gpurun_out/r02_warp $B > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_ont_launches.csv python bench.py --lean --workload ont --reads 300000 --steps 2 --warmup 3 --e2e-steps 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:amp_trim_pileup_ont -s 3 -c 1 -f -o gpurun_out/r02_ont python bench.py --lean --workload ont --reads 300000 --steps 2 --warmup 3 --e2e-steps 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:amp_bgzf_inflate -s 1 -c 1 -f -o gpurun_out/r02_inflate python profiles/tools/bam_step.py 1000000 3 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_write_launches.csv python profiles/tools/prof_write.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:amp_bgzf_deflate_kernel -s 1 -c 1 -f -o gpurun_out/r02_deflate python profiles/tools/prof_write.py > /dev/null 2>&1
for r in warp ont inflate deflate; do
  ncu -i gpurun_out/r02_$r.ncu-rep --page raw --csv > gpurun_out/r02_${r}_raw.csv 2>/dev/null
  ncu -i gpurun_out/r02_$r.ncu-rep --page details > gpurun_out/r02_${r}_details.txt 2>/dev/null
done
ls -la gpurun_out/r02_launches.csv gpurun_out/r02_ont_launches.csv gpurun_out/r02_*raw.csv
