# Round-2 multi-GPU check: N-GPU parity tests + bench with the parity gate at N ranks (N = number of visible GPUs)
set -x
N=$(nvidia-smi -L | wc -l)
TAG=${TAG:-r2_v1}
timeout 900 python -m pytest tests/test_multi_rank.py -m gpu -q 2>&1 | tail -15
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps ${STEPS:-10} --warmup 3 ${EXTRA} > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err; echo rc=$?
grep -v "^\[W\|^W1\|Warning" gpurun_out/${TAG}_bench_${N}gpu.err | tail -12 | cut -c1-1200; cut -c1-2500 gpurun_out/${TAG}_bench_${N}gpu.json
