#!/bin/bash
# round 2, session 18: tcgen05 attention with 8 softmax warps per CTA (vs 4), correctness sweep first, under short timeouts
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r2s18
timeout 120 python scripts/attn_debug.py > $O.attn_debug.log 2>&1; echo "attn_debug rc=$? nonfinite-lines: $(grep -c 'non-finite [1-9]' $O.attn_debug.log) bad-rows: $(grep -c 'rows>tol' $O.attn_debug.log)"; grep "rows>tol" $O.attn_debug.log | head -5
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_encoder.py -x -q -k "attention or encoder" > $O.tests_a.log 2>&1; echo "tests_a rc=$?"; tail -3 $O.tests_a.log
for sw in 8 4; do
  LB2_ATTN_WARPS=$sw timeout 200 python scripts/kernel_bench.py > $O.kb_sw$sw.log 2>&1; echo "== softmax warps=$sw rc=$?"; grep "attention" $O.kb_sw$sw.log
done
