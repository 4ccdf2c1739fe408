#!/bin/bash
# round 2, session 23 (1 GPU, the last seconds of the round's budget): the GPU test files session 22 did not reach
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r2s23
timeout 110 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_pq_pruning.py tests/test_gpu_vamana.py -x -q --durations=5 > $O.tests.log 2>&1; echo "tests rc=$?"; tail -8 $O.tests.log
