#!/bin/bash
# round 2, session 16: depth sensitivity of the weight-stationary GEMM (2 / 3 / 4 activation stages)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r2s16
for st in 4 3 2; do
  LB2_GEMM_WS_STAGES=$st timeout 200 python scripts/kernel_bench.py > $O.kb_st$st.log 2>&1; echo "== stages=$st rc=$?"; grep "gemm qkv\|gemm ffn-up" $O.kb_st$st.log
  LB2_GEMM_EXP=1 LB2_GEMM_WS_STAGES=$st timeout 200 python scripts/kernel_bench.py > $O.kb_st${st}_nostore.log 2>&1; echo "== stages=$st no stores"; grep "gemm qkv\|gemm ffn-up" $O.kb_st${st}_nostore.log
done
