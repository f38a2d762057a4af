# A/B of one environment switch on the SGBM micro-benchmark (33 KITTI frames), after the parity tests selected by $TESTS with the switch on.
# usage: VAR=SSM_X VALS="0 1" [TESTS="-k expr"] bash scripts/gpu_envab.sh
export $VAR=${TESTVAL:-1}
timeout 240 python -m pytest tests/test_gpu_sgbm.py -m gpu -x -q ${TESTS:-} 2>&1 | tail -3
unset $VAR
for v in ${VALS:-0 1}; do
  env $VAR=$v timeout 120 python scripts/prof_sgbm.py --batch 33 --reps 10 2>&1 | tail -1 | cut -c1-420
done
