# 8-GPU run: many-rank parity tests, bench configs 1 / 2 / 4 / 3 with the parity gate
set -x
N=$(nvidia-smi -L | wc -l)
TAG=${TAG:-r2_v1}
timeout 600 python -m pytest tests/test_multi_rank.py -m gpu -q -k many_gpu 2>&1 | tail -5
for C in ${CONFIGS:-1 2 4 3}; do
  ST=""; if [ "$C" = "1" ]; then ST="--steps 20"; fi
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$C bench.py --gpus $N --config $C $ST ${EXTRA} > gpurun_out/${TAG}_config${C}_${N}gpu.json 2> gpurun_out/${TAG}_config${C}_${N}gpu.err; echo rc=$?
  grep "parity:\|Error\|error" gpurun_out/${TAG}_config${C}_${N}gpu.err | tail -4 | cut -c1-700; cut -c1-1800 gpurun_out/${TAG}_config${C}_${N}gpu.json
done
