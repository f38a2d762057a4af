set -x
timeout 600 python -m pytest tests/test_gpu_sgbm.py -m gpu -x -q -k "cityscapes" 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()"
python - <<'PY'
# Cityscapes-shaped throughput (configs[3]): 2048x1024, 256 disparities, 19 classes, B = 4
import time, numpy as np, torch
from semantic_slam_mapping_b200 import Context, Params, synth
H, W, D, B = 1024, 2048, 256, 4
from semantic_slam_mapping_b200.params import cityscapes_params
p = cityscapes_params(max_batch=B, map_capacity=1 << 22, resolution=0.05)
seq = synth.sequence(B, H, W, D, 19, seed=3, distinct=2)
dev = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in seq.items() if k != "label"}
with Context(p) as ctx:
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(2):
        ctx.pipeline_batch_device(dev["left"], dev["right"], dev["semantic"], dev["rgb"], dev["pose"], B, W, H, stream=st)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        ctx.pipeline_batch_device(dev["left"], dev["right"], dev["semantic"], dev["rgb"], dev["pose"], B, W, H, stream=st)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("cityscapes 2048x1024 D=256: %.1f frames/s (B=%d), voxels %d" % (5 * B / dt, B, ctx.map_size()))
PY
