python -m pytest tests/test_gpu_sgbm.py tests/test_gpu_mapper.py -m gpu -x -q 2>&1 | tail -3
for b in 66 99; do
python bench.py --batch $b --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ab_b$b.json 2> gpurun_out/ab_b$b.err
python - <<PY
import json
d=json.load(open('gpurun_out/ab_b$b.json'))
print('batch $b', round(d['value'],1),'fps e2e',round(d['e2e']['value'],1), {k:round(v['ms_per_step'],3) for k,v in d['roofline']['stages'].items()}, d['clocks']['sm_mhz'], d['clocks']['reasons'], d['parity']['ok'])
PY
done
