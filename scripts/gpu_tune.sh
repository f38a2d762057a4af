# usage: CONFIGS="0,0 2,0 0,1" bash scripts/gpu_tune.sh   -> per-kernel time (ncu, B=33) for each SSM_TUNE0,SSM_TUNE1[,2,3]
for cfg in ${CONFIGS:-0,0}; do
  IFS=, read t0 t1 t2 t3 <<< "$cfg"
  export SSM_TUNE0=${t0:-0} SSM_TUNE1=${t1:-0} SSM_TUNE2=${t2:-0} SSM_TUNE3=${t3:-0}
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --batch ${B:-33} --input-batches 3 > gpurun_out/tune.json 2> gpurun_out/tune.err
  python - <<PY
import json
d=json.load(open('gpurun_out/tune.json'))
print('cfg=$cfg', round(d['value'],1),'fps', {k:v['ms_per_step'] for k,v in d['roofline']['stages'].items()})
PY
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_vertical3|k_hsweep|k_cost_fused|k_hfwd|k_hrev" -s 8 -c 4 --csv --log-file gpurun_out/tune_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --batch ${B:-33} --input-batches 2 > /dev/null 2>&1
  python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/tune_launches.csv')) if len(r)>10 and r[0].isdigit()]
print('   ', '  '.join(f"{r[4][:18]} {int(r[-1])/1e3:.0f}us" for r in rows))
PY
done
