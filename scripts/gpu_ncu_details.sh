# ncu full-set capture of the main kernels + the details page (scheduler / warp-state sections) as text
set -x
ncu --set full --clock-control none --import-source on -k regex:"${KREGEX:-k_cost_fused|k_vertical3|k_hrev|k_hfwd}" -s ${SKIP:-4} -c ${COUNT:-4} -o gpurun_out/${OUT:-details} -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --batch 33 --input-batches 1 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ncu -i gpurun_out/${OUT:-details}.ncu-rep --page details --section SchedulerStats --section WarpStateStats --section SpeedOfLight --section Occupancy > gpurun_out/${OUT:-details}_details.txt 2>&1
