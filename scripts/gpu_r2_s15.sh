#!/bin/bash
# round 2, session 15: weight-stationary GEMM epilogue with direct global stores instead of TMA stores (LB2_GEMM_EXP=2)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r2s15
for e in 0 2; do
  LB2_GEMM_EXP=$e timeout 200 python scripts/kernel_bench.py > $O.kb_exp$e.log 2>&1; echo "== exp=$e rc=$?"; grep "gemm qkv\|gemm ffn-up" $O.kb_exp$e.log
done
LB2_GEMM_EXP=2 timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k gemm > $O.tests_exp2.log 2>&1; echo "tests rc=$?"; tail -3 $O.tests_exp2.log
