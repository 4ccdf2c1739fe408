#!/bin/bash
# round 2, session 22 (1 GPU, last call of the round, <= 2.5 min): ncu --set full of the tcgen05 attention kernel at bench size
# (the r02d capture was a smoke-sized launch), then as many of the GPU parity tests as fit
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r2s22
timeout 75 ncu --set full --clock-control none --import-source on -k regex:attention_tc_kernel -s 2 -c 1 -o $O.attn \
    python scripts/attn_profile_target.py 3 > $O.ncu.log 2>&1; echo "ncu rc=$?"; tail -2 $O.ncu.log
timeout 80 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_search.py tests/test_gpu_kernels.py tests/test_gpu_pq_pruning.py tests/test_gpu_vamana.py -x -q --durations=5 > $O.tests.log 2>&1; echo "tests rc=$?"; tail -3 $O.tests.log
