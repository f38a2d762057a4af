# Round-2 check run (1 GPU): GPU tests, smoke, bench with the parity gate (short), config 3 and 4 sanity.
set -x
TAG=${TAG:-r2_v1}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo rc=$?; tail -c 1500 gpurun_out/${TAG}_bench.err; cut -c1-3000 gpurun_out/${TAG}_bench.json
