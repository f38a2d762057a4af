# Round-end evidence run (1 GPU): tests, bench (ours + reference arm), ncu launch list, ncu full of the main kernels.
set -x
TAG=${TAG:-r1_final}
python -m pytest tests -m gpu -q 2>&1 | tail -3
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 400 gpurun_out/${TAG}_bench.err; cut -c1-600 gpurun_out/${TAG}_bench.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>/dev/null; cut -c1-300 gpurun_out/${TAG}_bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 80 --csv --log-file gpurun_out/${TAG}_launches_ncu.csv python bench.py --steps 3 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_prefilter_tab|k_cost_tma|k_vertical3|k_hfwd|k_hrev|k_points_fuse|k_select_fused|k_cc_apply_bands" -s 9 -c 9 -o gpurun_out/${TAG}_top -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
python scripts/bench_rows.py > gpurun_out/${TAG}_rows.json 2> gpurun_out/rows.err; tail -2 gpurun_out/rows.err; cat gpurun_out/${TAG}_rows.json
# optional: another BASELINE config on the same box (CONFIGS="3 4")
for k in ${CONFIGS:-}; do
  python bench.py --config $k --no-cpu-baseline > gpurun_out/${TAG}_config${k}_1gpu.json 2> gpurun_out/${TAG}_config${k}_1gpu.err; cut -c1-400 gpurun_out/${TAG}_config${k}_1gpu.json
done
