#!/bin/bash
# round 2, session 2: attention micro-benchmark + ncu, PQ pruning tests, graph recall variants at 10 M
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r2s2
python -c "from leann_b200 import build; build.build()" > $O.build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_graph_build.py tests/test_gpu_pq_pruning.py -x -q -s > $O.tests.log 2>&1; echo "tests rc=$?"; grep -v "^$" $O.tests.log | tail -12
timeout 300 python scripts/kernel_bench.py > $O.kb_tc.log 2>&1; grep attention $O.kb_tc.log
LB2_ATTN_LEGACY=1 timeout 300 python scripts/kernel_bench.py > $O.kb_legacy.log 2>&1; grep attention $O.kb_legacy.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tc_kernel -c 2 -o $O.attn_tc python scripts/kernel_bench.py 131072 > $O.ncu_attn.log 2>&1; echo "ncu rc=$?"
timeout 1500 python scripts/graph_recall_10m.py 10000000 > $O.graph_recall.log 2>&1; echo "graph rc=$?"; grep -v "^$" $O.graph_recall.log | tail -30
