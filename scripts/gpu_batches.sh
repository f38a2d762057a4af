# frames/s for several batch sizes / sub-batch stream counts (1 GPU)
run() {  # name, batch, env...
  name=$1; b=$2; shift 2
  env "$@" python bench.py --steps ${STEPS:-12} --warmup 3 --no-cpu-baseline --batch $b > gpurun_out/bt_$name.json 2> gpurun_out/bt_$name.err || tail -3 gpurun_out/bt_$name.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bt_$name.json'))
print('$name', round(d['value'],1),'fps e2e',round(d['e2e']['value'],1), 'ms/step', round(d['ms_per_step'],3))
PY
}
run b66 66 A=1
run b99 99 A=1
run b132_s4 132 SSM_TUNE3=4
run b132_s3 132 SSM_TUNE3=3
run b66_s3 66 SSM_TUNE3=3
run b33_s1 33 A=1
