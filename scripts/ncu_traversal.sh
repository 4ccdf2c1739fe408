#!/bin/bash
# ncu --set full of the two traversal kernels (second launch of each = warm), 1 GPU.
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"hnsw_step_kernel|vamana_search_kernel" -c 4 -o gpurun_out/traversal \
    python scripts/traversal_profile.py 400000 8192 > gpurun_out/ncu_traversal.log 2>&1
echo "ncu traversal rc=$?"; tail -4 gpurun_out/ncu_traversal.log
