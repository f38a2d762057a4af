set -x
N=${N:-2}
nvidia-smi -L
timeout 600 python -m pytest tests/test_multi_rank.py -m gpu -x -q 2>&1 | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --batch 33 --input-batches 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -c 1800 gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
