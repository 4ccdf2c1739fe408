#!/bin/bash
# round 2, session 6a: attention v2.1, PQ pruning tests, corpus variants for the 10 M recall target
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r2s6a
python -c "from leann_b200 import build; build.build(force=True)" > $O.build.log 2>&1
timeout 300 python scripts/attn_debug.py > $O.attn_debug.log 2>&1; echo "attn_debug rc=$?"; grep -c "non-finite 0 " $O.attn_debug.log; grep "non-finite [1-9]" $O.attn_debug.log | head -5
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "attention or res_ln" > $O.test_kernels.log 2>&1; arc=$?; echo "test_kernels rc=$arc"; tail -3 $O.test_kernels.log
timeout 300 python scripts/kernel_bench.py > $O.kb.log 2>&1; grep "attention" $O.kb.log
if [ $arc -ne 0 ]; then export LB2_ATTN_LEGACY=1; echo "USING LEGACY ATTENTION for the rest"; fi
timeout 600 python -m pytest tests/test_gpu_pq_pruning.py tests/test_gpu_encoder.py -x -q -s > $O.tests.log 2>&1; echo "tests rc=$?"; grep -v "^$" $O.tests.log | grep "max |d\|passed\|failed\|Error" | tail -12
LB2_P_TOPIC=0.85 LB2_P_SUPER=0.08 timeout 600 python scripts/graph_recall_10m.py 10000000 base: > $O.graph_p85.log 2>&1; echo "p85 rc=$?"; grep -v "^$" $O.graph_p85.log | tail -4
LB2_TOPIC_SIZE=64 timeout 600 python scripts/graph_recall_10m.py 10000000 base: > $O.graph_t64.log 2>&1; echo "t64 rc=$?"; grep -v "^$" $O.graph_t64.log | tail -4
LB2_P_TOPIC=0.90 LB2_P_SUPER=0.05 timeout 600 python scripts/graph_recall_10m.py 10000000 base: > $O.graph_p90.log 2>&1; echo "p90 rc=$?"; grep -v "^$" $O.graph_p90.log | tail -4
