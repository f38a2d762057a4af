set -x
ncu --set full --clock-control none --import-source on -k regex:"${KREGEX:-k_cost_fused|k_vertical3|k_hsweep}" -s ${SKIP:-3} -c ${COUNT:-3} -o gpurun_out/${OUT:-prof} -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --batch ${B:-33} --input-batches 1 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
