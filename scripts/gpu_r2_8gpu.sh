# 8-GPU run: many-rank parity tests, bench configs (default 1 3 4) with the parity gate, NVLink byte counters around config 1
set -x
N=$(nvidia-smi -L | wc -l)
TAG=${TAG:-r2_v2}
timeout 600 python -m pytest tests/test_multi_rank.py -m gpu -q -k many_gpu 2>&1 | tail -3
for C in ${CONFIGS:-1 3 4}; do
  ST=""; if [ "$C" = "1" ]; then ST="--steps 20"; nvidia-smi nvlink -gt d -i 0 > gpurun_out/${TAG}_nvlink_before.txt 2>&1; fi
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$C bench.py --gpus $N --config $C $ST ${EXTRA} > gpurun_out/${TAG}_config${C}_${N}gpu.json 2> gpurun_out/${TAG}_config${C}_${N}gpu.err; echo rc=$?
  if [ "$C" = "1" ]; then nvidia-smi nvlink -gt d -i 0 > gpurun_out/${TAG}_nvlink_after.txt 2>&1; fi
  grep "parity:\|Error\|error" gpurun_out/${TAG}_config${C}_${N}gpu.err | tail -4 | cut -c1-500; cut -c1-400 gpurun_out/${TAG}_config${C}_${N}gpu.json
done
