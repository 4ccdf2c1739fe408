#!/bin/bash
# round 2, session 21 (1 GPU, short): attention tail-block rotation — kernel tests, then the A/B timing
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r2s21
timeout 200 python -m pytest tests/test_gpu_kernels.py -x -q -k "attention" > $O.tests.log 2>&1; echo "tests rc=$?"; tail -3 $O.tests.log
timeout 120 python scripts/attn_rotate_ab.py > $O.ab.log 2>&1; echo "ab rc=$?"; cat $O.ab.log | tail -8
