#!/bin/bash
# round 2, session 11: tcgen05.mma issue-rate microbenchmark (cta_group 1 vs 2, with background smem traffic), ncu of the
# CTA-pair LN GEMM, kernel microbench with the pair kernel as default
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r2s11
python -c "from leann_b200 import build; build.needs_build() and build.build()" > $O.build.log 2>&1
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I leann_b200/csrc scripts/umma_rate.cu -o /tmp/umma_rate > $O.umma_build.log 2>&1
timeout 120 /tmp/umma_rate > $O.umma_rate.log 2>&1; echo "umma rc=$?"; cat $O.umma_rate.log
timeout 300 python scripts/kernel_bench.py > $O.kb.log 2>&1; echo "== kb rc=$?"; grep "gemm" $O.kb.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_f16_ln_pair_kernel -c 2 -o $O.gemm_ln_pair python scripts/kernel_bench.py > $O.ncu_pair.log 2>&1; echo "ncu rc=$?"
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_encoder.py -x -q > $O.tests_a.log 2>&1; echo "tests_a rc=$?"; tail -2 $O.tests_a.log
