#!/bin/bash
# round 2, session 14: what the weight-stationary GEMM waits for — timing with the TMA stores of the epilogue switched off
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r2s14
for e in 0 1; do
  LB2_GEMM_EXP=$e timeout 200 python scripts/kernel_bench.py > $O.kb_exp$e.log 2>&1; echo "== exp=$e rc=$?"; grep "gemm qkv\|gemm ffn-up" $O.kb_exp$e.log
done
