set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for B in ${BATCHES:-20 33}; do
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --batch $B --input-batches 3 > gpurun_out/bench_quick_$B.json 2> gpurun_out/bench_quick.err; python - <<PY
import json
d=json.load(open('gpurun_out/bench_quick_$B.json'))
print('B=$B', round(d['value'],1),'fps e2e',round(d['e2e']['value'],1), {k:v['ms_per_step'] for k,v in d['roofline']['stages'].items()})
PY
tail -3 gpurun_out/bench_quick.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 24 --csv --log-file gpurun_out/quick_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --batch 33 --input-batches 2 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/quick_launches.csv')) if len(r)>10 and r[0].isdigit()]
for r in rows[:24]: print(r[4][:60].ljust(60), r[8].ljust(18), int(r[-1])/1e3, 'us')
PY
