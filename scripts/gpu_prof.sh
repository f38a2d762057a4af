# one ncu --set full capture of the kernels matching $KREGEX from the SGBM micro-benchmark (B = 33 KITTI frames)
ncu --set full --clock-control none --import-source on -k regex:"${KREGEX:-k_cost_tma}" -s ${SKIP:-1} -c ${COUNT:-1} -o gpurun_out/${OUT:-prof} -f python scripts/prof_sgbm.py --reps 2 ${ARGS:-} > gpurun_out/${OUT:-prof}.log 2>&1
tail -3 gpurun_out/${OUT:-prof}.log
