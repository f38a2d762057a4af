# vertical-kernel variants (SSM_TUNE2): correctness on the SGBM tests, then stage times at B = 33
set -x
python -m pytest tests/test_gpu_sgbm.py -x -q 2>&1 | tail -3
for T in ${TUNES:-0 2 3 9}; do
SSM_TUNE2=$T python bench.py --steps 8 --warmup 3 --no-cpu-baseline --batch 33 --input-batches 2 > gpurun_out/tune2_$T.json 2> gpurun_out/tune2.err; python - <<PY
import json
d=json.load(open('gpurun_out/tune2_$T.json'))
print('TUNE2=$T', round(d['value'],1),'fps', {k:v['ms_per_step'] for k,v in d['roofline']['stages'].items()})
PY
done
