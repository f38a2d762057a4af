# Vertical cluster kernel: co-resident clusters and the time of one wave for cluster sizes 4..8 (KITTI / 128).
# usage: bash scripts/gpu_clusters.sh   (writes gpurun_out/clusters.txt)
python -m pytest tests/test_gpu_sgbm.py -m gpu -x -q -k "all_cluster_sizes" 2>&1 | tail -3
for cs in 4 5 6 7 8; do
  for b in ${BATCHES:-16}; do
    echo "== cluster $cs batch $b"
    SSM_DEBUG_CLUSTERS=1 SSM_MIN_CLUSTER=$cs python scripts/prof_sgbm.py --batch $b --reps 5 2>&1 | grep -v "^$" | sort -u | tail -4
  done
done 2>&1 | tee gpurun_out/clusters.txt
