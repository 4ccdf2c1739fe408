#!/bin/bash
# round 2, session 3: attention v2 + fused LN GEMM + graph recall with the fixed-size super-topic corpus at 10 M, PQ tests
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r2s3
python -c "from leann_b200 import build; build.build()" > $O.build.log 2>&1
timeout 300 python scripts/attn_debug.py > $O.attn_debug.log 2>&1; echo "attn_debug rc=$?"; grep "==" $O.attn_debug.log | awk '{print $(NF-5), $0}' | sort -g | tail -3
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "attention or res_ln" > $O.test_kernels.log 2>&1; arc=$?; echo "test_kernels rc=$arc"; tail -3 $O.test_kernels.log
timeout 300 python scripts/kernel_bench.py > $O.kb.log 2>&1; grep "attention\|ln\|layernorm\|attn-out\|ffn-down" $O.kb.log
if [ $arc -ne 0 ]; then export LB2_ATTN_LEGACY=1; echo "USING LEGACY ATTENTION for the rest"; fi
timeout 900 python -m pytest tests/test_gpu_graph_build.py tests/test_gpu_pq_pruning.py -x -q -s > $O.tests.log 2>&1; echo "tests rc=$?"; grep -v "^$" $O.tests.log | tail -12
timeout 900 python scripts/graph_recall_10m.py 10000000 base: sweep1:sweeps=1 > $O.graph_recall.log 2>&1; echo "graph rc=$?"; grep -v "^$" $O.graph_recall.log | tail -12
