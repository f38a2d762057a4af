# A/B of one environment switch on the 1-GPU bench (stage times + frames/s), after the GPU test suite.
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
run() {  # name, env...
  name=$1; shift
  env "$@" python bench.py --steps ${STEPS:-20} --warmup 3 --no-cpu-baseline > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
  python - <<PY
import json
d=json.load(open('gpurun_out/ab_$name.json'))
print('$name', round(d['value'],1),'fps e2e',round(d['e2e']['value'],1), {k:round(v['ms_per_step'],3) for k,v in d['roofline']['stages'].items()}, d['clocks'])
PY
}
run new A=1
run legacy_select SSM_LEGACY_SELECT=1
for extra in ${EXTRA:-}; do run "x_${extra//=/_}" $extra; done
