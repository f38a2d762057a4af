# BASELINE configs 2, 3, 4 through bench.py on the visible GPUs
set -x
N=$(nvidia-smi -L | wc -l)
TAG=${TAG:-r2_v1}
for C in ${CONFIGS:-3 2 4}; do
  if [ "$N" = "1" ]; then
    timeout 1200 python bench.py --config $C --no-cpu-baseline ${EXTRA} > gpurun_out/${TAG}_config${C}_${N}gpu.json 2> gpurun_out/${TAG}_config${C}_${N}gpu.err; echo rc=$?
  else
    timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$C bench.py --gpus $N --config $C ${EXTRA} > gpurun_out/${TAG}_config${C}_${N}gpu.json 2> gpurun_out/${TAG}_config${C}_${N}gpu.err; echo rc=$?
  fi
  grep -v "^\[W\|^W1\|Warning\|^\*\*\*\|OMP_NUM" gpurun_out/${TAG}_config${C}_${N}gpu.err | tail -6 | cut -c1-900; cut -c1-3500 gpurun_out/${TAG}_config${C}_${N}gpu.json
done
