#!/bin/bash
# round 2, session 7: attention with empty-warp skip (+ncu), level-1 coverage promotion on the topic-size-64 corpus at 10 M
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r2s7
python -c "from leann_b200 import build; build.build(force=True)" > $O.build.log 2>&1
timeout 300 python scripts/attn_debug.py > $O.attn_debug.log 2>&1; echo "attn_debug rc=$?"; grep -c "non-finite 0 " $O.attn_debug.log; grep "non-finite [1-9]" $O.attn_debug.log | head -5
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "attention" > $O.test_kernels.log 2>&1; echo "test_kernels rc=$?"; tail -2 $O.test_kernels.log
timeout 300 python scripts/kernel_bench.py > $O.kb.log 2>&1; grep "attention" $O.kb.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_tc_kernel -s 13 -c 1 -o $O.attn_tc python scripts/kernel_bench.py > $O.ncu_attn.log 2>&1; echo "ncu rc=$?"
LB2_TOPIC_SIZE=64 LB2_BENCH_VERBOSE=1 timeout 1200 python scripts/graph_recall_10m.py 10000000 cover1:cover=1 cover2:cover=2 cov1sw1:cover=1,sweeps=1 > $O.graph_t64.log 2>&1; echo "t64 rc=$?"; grep -v "^$" $O.graph_t64.log | grep -v inserted | tail -14
LB2_TOPIC_SIZE=64 LB2_P_TOPIC=0.90 LB2_P_SUPER=0.05 timeout 900 python scripts/graph_recall_10m.py 10000000 base: cover1:cover=1 > $O.graph_p90t64.log 2>&1; echo "p90t64 rc=$?"; grep -v "^$" $O.graph_p90t64.log | tail -8
