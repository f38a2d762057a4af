# one ncu --set full capture of the kernels matching $KREGEX (default: the fused selection kernel), B = 33
set -x
ncu --set full --clock-control none --import-source on -k regex:"${KREGEX:-k_select_fused}" -s ${SKIP:-1} -c ${COUNT:-1} -o gpurun_out/${OUT:-sel} -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --batch ${B:-33} --input-batches 1 > gpurun_out/ncu_one.log 2>&1
tail -3 gpurun_out/ncu_one.log
