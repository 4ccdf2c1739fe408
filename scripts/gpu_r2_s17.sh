#!/bin/bash
# round 2, session 17: does the epilogue's tcgen05.ld traffic slow the MMAs down? (umma_rate with background TMEM reads)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r2s17
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I leann_b200/csrc scripts/umma_rate.cu -o /tmp/umma_rate > $O.umma_build.log 2>&1
timeout 120 /tmp/umma_rate > $O.umma_rate.log 2>&1; echo "umma rc=$?"; cat $O.umma_rate.log
