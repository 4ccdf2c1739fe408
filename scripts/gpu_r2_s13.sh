#!/bin/bash
# round 2, session 13: TMA load throughput microbenchmark
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r2s13
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I leann_b200/csrc scripts/tma_rate.cu -o /tmp/tma_rate -lcuda > $O.build.log 2>&1
timeout 120 /tmp/tma_rate 64 > $O.tma_l2.log 2>&1; echo "rc=$?"; cat $O.tma_l2.log
timeout 120 /tmp/tma_rate 4096 > $O.tma_hbm.log 2>&1; echo "rc=$?"; cat $O.tma_hbm.log
