#!/bin/bash
# round 2, session 5: attention v2 (really), fused LN, PQ pruning, graph recall on the fixed-size super-topic corpus, first full 10 M bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r2s5
grep -c "h2exp2" leann_b200/csrc/attention_tc.cu
python -c "from leann_b200 import build; build.build(force=True)" > $O.build.log 2>&1
timeout 300 python scripts/attn_debug.py > $O.attn_debug.log 2>&1; echo "attn_debug rc=$?"; grep -c "non-finite 0 " $O.attn_debug.log; grep "non-finite [1-9]" $O.attn_debug.log | head -5
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "attention or res_ln" > $O.test_kernels.log 2>&1; arc=$?; echo "test_kernels rc=$arc"; tail -3 $O.test_kernels.log
timeout 300 python scripts/kernel_bench.py > $O.kb.log 2>&1; grep "attention" $O.kb.log
if [ $arc -ne 0 ]; then export LB2_ATTN_LEGACY=1; echo "USING LEGACY ATTENTION for the rest"; fi
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_tc_kernel -s 2 -c 1 -o $O.attn_tc python scripts/kernel_bench.py > $O.ncu_attn.log 2>&1; echo "ncu rc=$?"
timeout 600 python -m pytest tests/test_gpu_pq_pruning.py tests/test_gpu_encoder.py -x -q -s > $O.tests.log 2>&1; echo "tests rc=$?"; grep -v "^$" $O.tests.log | grep "max |d\|passed\|failed\|Error" | tail -12
timeout 600 python scripts/graph_recall_10m.py 10000000 base: > $O.graph_recall.log 2>&1; echo "graph rc=$?"; grep -v "^$" $O.graph_recall.log | tail -6
timeout 1200 python bench.py --steps 5 --warmup 3 --budget-s 100 > $O.bench_10m.json 2> $O.bench_10m.err; echo "bench10m rc=$?"
grep "\[bench\]" $O.bench_10m.err | tail -12; python -c "
import json;d=json.load(open('$O.bench_10m.json'));print({k:d[k] for k in ('value','recall_at_10')}, d['e2e']['value'], d['roofline']['frac'], {k:d['detail'][k] for k in ('attention_share','layernorm_share','encoder_algorithmic_tflops','ndis_per_query')}, d.get('cpu_baseline'), d.get('parity'))"
