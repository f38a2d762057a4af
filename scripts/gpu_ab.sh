# A/B of environment switches on the 1-GPU bench (stage times + frames/s), after the GPU test suite.
# usage: [TESTS="-k expr"] [EXTRA="A=1 B=2 ..."] [NCU=regex] bash scripts/gpu_ab.sh
set -x
python -m pytest tests -m gpu -x -q ${TESTS:-} 2>&1 | tail -6
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
for extra in ${EXTRA:-}; do run "x_${extra//=/_}" $extra; done
if [ -n "${NCU:-}" ]; then
ncu --set full --clock-control none --import-source on -k regex:"$NCU" -s ${SKIP:-1} -c ${COUNT:-1} -o gpurun_out/${OUT:-sel} -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --batch 33 --input-batches 1 > gpurun_out/ncu_one.log 2>&1
tail -2 gpurun_out/ncu_one.log
fi
