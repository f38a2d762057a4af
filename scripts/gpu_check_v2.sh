set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1_v2.json 2> gpurun_out/bench_r1_v2.err; tail -c 3000 gpurun_out/bench_r1_v2.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r1_v2.json 2>&1; tail -c 1500 gpurun_out/bench_ref_r1_v2.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 60 --csv --log-file gpurun_out/r1_v2_launches_ncu.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_hsweep|k_aggr_path|k_pix_hsum|k_vsum" -s 4 -c 6 -o gpurun_out/r1_v2_top python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
