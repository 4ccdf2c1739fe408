#!/bin/bash
# round 2, session 10: warp-uniform tcgen05 issue path (GEMMs, attention), CTA-pair LN GEMM (tests under a short timeout), short 10 M bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r2s10
python -c "from leann_b200 import build; build.needs_build() and build.build()" > $O.build.log 2>&1
timeout 300 python scripts/kernel_bench.py > $O.kb.log 2>&1; echo "== kb rc=$?"; grep "gemm\|attention" $O.kb.log
LB2_GEMM_WS=0 timeout 300 python scripts/kernel_bench.py > $O.kb_ws0.log 2>&1; echo "== ws0"; grep "gemm" $O.kb_ws0.log | head -4
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_encoder.py -x -q -k "not cta_pair" > $O.tests_a.log 2>&1; echo "tests_a rc=$?"; tail -2 $O.tests_a.log
timeout 300 python scripts/attn_debug.py > $O.attn_debug.log 2>&1; echo "attn_debug rc=$? nonfinite-lines: $(grep -c 'non-finite [1-9]' $O.attn_debug.log) bad-rows: $(grep -c 'rows>tol' $O.attn_debug.log)"
timeout 180 python -m pytest tests/test_gpu_kernels.py -x -q -k "cta_pair" > $O.tests_pair.log 2>&1; echo "tests_pair rc=$?"; tail -4 $O.tests_pair.log
LB2_GEMM_LN_PAIR=1 timeout 180 python scripts/kernel_bench.py > $O.kb_pair.log 2>&1; echo "== pair rc=$?"; grep "fused" $O.kb_pair.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_f16_ws_kernel -c 1 -o $O.gemm_ws python scripts/kernel_bench.py > $O.ncu_ws.log 2>&1; echo "ncu rc=$?"
timeout 900 python bench.py --gpus 1 --steps 6 --warmup 3 --budget-s 90 > $O.bench.json 2> $O.bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('$O.bench.json'))
print({k:d.get(k) for k in ('value','recall_at_10','ms_per_step','gpu_launches')}); print(d['e2e']['value']); print(d['roofline']); print(d.get('parity')); print({k:d['detail'][k] for k in ('encoder_share','attention_share','encoder_algorithmic_tflops')})
PY
t0=$(date +%s)
timeout 1800 python -m pytest tests -x -q -m gpu -k "not cta_pair" > $O.tests.log 2>&1; echo "tests rc=$? wall=$(( $(date +%s) - t0 ))s"; tail -5 $O.tests.log
