#!/bin/bash
# round 2, session 4: attention v2 validation + micro-benchmarks, PQ pruning under compute-sanitizer, graph recall on the
# fixed-size super-topic corpus at 10 M
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r2s4
git log --oneline 2>/dev/null | head -1
python -c "from leann_b200 import build; build.build(force=True)" > $O.build.log 2>&1
timeout 300 python scripts/attn_debug.py > $O.attn_debug.log 2>&1; echo "attn_debug rc=$?"; grep -c "non-finite 0 " $O.attn_debug.log; grep "non-finite [1-9]" $O.attn_debug.log | head -5
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "attention or res_ln" > $O.test_kernels.log 2>&1; arc=$?; echo "test_kernels rc=$arc"; tail -3 $O.test_kernels.log
timeout 300 python scripts/kernel_bench.py > $O.kb.log 2>&1; grep "attention\|fused\|layernorm\|attn-out\|ffn-down" $O.kb.log
if [ $arc -ne 0 ]; then export LB2_ATTN_LEGACY=1; echo "USING LEGACY ATTENTION for the rest"; fi
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_pq_pruning.py -x -q -k "oracle and mips" > $O.sanitizer.log 2>&1; echo "sanitizer rc=$?"; grep -A12 "Invalid\|ERROR SUMMARY" $O.sanitizer.log | head -50
timeout 600 python -m pytest tests/test_gpu_graph_build.py -x -q -s > $O.tests.log 2>&1; echo "tests rc=$?"; grep -v "^$" $O.tests.log | tail -8
timeout 900 python scripts/graph_recall_10m.py 10000000 base: sweep1:sweeps=1 > $O.graph_recall.log 2>&1; echo "graph rc=$?"; grep -v "^$" $O.graph_recall.log | tail -12
