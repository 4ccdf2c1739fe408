#!/bin/bash
# round 2, session 9: weight-stationary GEMM (tests, A/B, ncu), full GPU test suite, short 10 M bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r2s9
python -c "from leann_b200 import build; build.needs_build() and build.build()" > $O.build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k "gemm" > $O.tests_gemm.log 2>&1; echo "tests_gemm rc=$?"; tail -3 $O.tests_gemm.log
timeout 300 python scripts/kernel_bench.py > $O.kb_ws1.log 2>&1; echo "== ws1"; grep "gemm" $O.kb_ws1.log
LB2_GEMM_WS=0 timeout 300 python scripts/kernel_bench.py > $O.kb_ws0.log 2>&1; echo "== ws0"; grep "gemm" $O.kb_ws0.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_f16_ws_kernel -c 2 -o $O.gemm_ws python scripts/kernel_bench.py > $O.ncu_ws.log 2>&1; echo "ncu rc=$?"
t0=$(date +%s)
timeout 1800 python -m pytest tests -x -q -m gpu > $O.tests.log 2>&1; echo "tests rc=$? wall=$(( $(date +%s) - t0 ))s"; tail -5 $O.tests.log
timeout 900 python bench.py --gpus 1 --steps 6 --warmup 3 --budget-s 90 > $O.bench.json 2> $O.bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('$O.bench.json'))
print({k:d.get(k) for k in ('value','recall_at_10','ms_per_step','gpu_launches')}); print(d['e2e']['value']); print(d['roofline']); print(d.get('parity')); print({k:d['detail'][k] for k in ('encoder_share','attention_share','encoder_algorithmic_tflops')})
PY
